"""Geometry containers with the reference's names (crl_kino/utils/rigid.py).

Only the data side lives here: rectangles are float32 like the reference's
(rigid.py:146, :263); every distance / collision computation runs in the CUDA
kernels (crlk_valid_state).  `load_obstacles` reads the reference's obstacle
pickles (lists of crl_kino.utils.rigid.RectObs) without the reference installed.
"""
import io
import pickle
from abc import ABC

import numpy as np


class Obstacle(ABC):
    def __init__(self, color="black"):
        self.color = color


class RectObs(Obstacle):
    def __init__(self, rect, color="black"):
        """rect: [[left, bot], [right, top]]  (rigid.py:139-146)."""
        super().__init__(color=color)
        rect = np.asarray(rect)
        assert np.any(rect[1] > rect[0]), "Not a normal rectangle"
        self.rect = np.array(rect, dtype=np.float32)

    def flat(self):
        return self.rect.reshape(4)

    def point_in_obstacle(self, point):
        x, y = point
        return bool(self.rect[0, 0] <= x <= self.rect[1, 0] and self.rect[0, 1] <= y <= self.rect[1, 1])


class Rigid(ABC):
    def __init__(self, pose=np.zeros(3), color="black"):
        self.color = color
        self.pose = np.array(pose, dtype=float)


class RectRobot(Rigid):
    def __init__(self, rect, pose=np.zeros(3), color="black"):
        super().__init__(pose, color)
        self.rect = np.array(rect, dtype=np.float32)


class CircleRobot(Rigid):
    """rigid.py:331-347: circular footprint about the body origin."""

    def __init__(self, radius, pose=np.zeros(3), color="black"):
        super().__init__(pose, color)
        self.radius = radius


def rects_of(obs_list):
    """list[RectObs] (or an (M, 4) / (M, 2, 2) array) -> float32 [M, 4] left, bottom, right, top."""
    if isinstance(obs_list, np.ndarray):
        return np.ascontiguousarray(obs_list.astype(np.float32).reshape(-1, 4))
    if len(obs_list) == 0:
        return np.zeros((0, 4), dtype=np.float32)
    return np.ascontiguousarray(np.stack([np.asarray(o.rect, dtype=np.float32).reshape(4) for o in obs_list]))


class _RefUnpickler(pickle.Unpickler):
    """Maps the reference's class paths onto this package's containers."""

    def find_class(self, module, name):
        if module.startswith("crl_kino.utils.rigid"):
            return {"RectObs": RectObs, "RectRobot": RectRobot, "CircleRobot": CircleRobot}[name]
        return super().find_class(module, name)


def load_obstacles(path_or_bytes):
    """Read a reference obstacle pickle (data/obstacles/*.pkl)."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        return _RefUnpickler(io.BytesIO(path_or_bytes)).load()
    with open(path_or_bytes, "rb") as f:
        return _RefUnpickler(f).load()
