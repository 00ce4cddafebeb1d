"""euler2mat stub: the reference only calls euler2mat(0, 0, yaw); the real
transforms3d evaluates it with math.cos / math.sin, so for ai = aj = 0 the
rotation is exactly [[c, -s, 0], [s, c, 0], [0, 0, 1]]."""
import math
import numpy as np


def euler2mat(ai, aj, ak, axes='sxyz'):
    assert ai == 0 and aj == 0
    c, s = math.cos(ak), math.sin(ak)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
