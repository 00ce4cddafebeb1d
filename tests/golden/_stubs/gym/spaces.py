"""spaces.Box stub. float32 default dtype matters: the reference's action
bounds become float32 tensors inside its Actor."""
import numpy as np


class Box:
    def __init__(self, low=None, high=None, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            self.low = np.asarray(low).astype(self.dtype)
            self.high = np.asarray(high).astype(self.dtype)
            self.shape = self.low.shape
        else:
            self.shape = tuple(shape)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)
