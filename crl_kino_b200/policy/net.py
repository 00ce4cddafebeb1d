"""Actor container (reference: crl_kino/policy/net.py:114-154).  Holds the MLP
weights; the forward pass runs in crlk_policy_forward."""
import numpy as np

from .. import ops


class Actor:
    def __init__(self, layer, state_shape, action_shape, action_range, device="cuda", seed=None):
        import torch
        self.layer = list(layer)
        dims = [int(np.prod(state_shape))] + self.layer + [int(np.prod(action_shape))]
        self.dims = dims
        if seed is not None:
            torch.manual_seed(seed)
        # same initialisation (and RNG consumption) as torch.nn.Linear on the CPU
        lins = [torch.nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)]
        self.weights = [l.weight.detach().numpy().copy() for l in lins]
        self.biases = [l.bias.detach().numpy().copy() for l in lins]
        self.low = np.asarray(action_range[0], dtype=np.float32)
        self.high = np.asarray(action_range[1], dtype=np.float32)
        self._handle = None

    def load_state_dict(self, sd, prefix="model."):
        """Accepts the reference's `actor.model.{0,2,4,...}.{weight,bias}` tensors."""
        keys = sorted({int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix)})
        assert len(keys) == len(self.weights), "layer count mismatch"
        for i, k in enumerate(keys):
            self.weights[i] = np.asarray(sd[f"{prefix}{k}.weight"], dtype=np.float32).copy()
            self.biases[i] = np.asarray(sd[f"{prefix}{k}.bias"], dtype=np.float32).copy()
        self._handle = None

    @property
    def handle(self):
        if self._handle is None:
            self._handle = ops.PolicyHandle(self.weights, self.biases, self.low, self.high)
        return self._handle
