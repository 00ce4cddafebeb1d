"""crlk_nearest vs RRT.choose_parent goldens and the oracle (float64 and float32-screened scans)."""
import numpy as np
import pytest
import torch

from conftest import golden
from gpu_util import np_
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _run(xy, targets, f32, n_nodes=None, stride=None):
    from crl_kino_b200 import ops
    xy = np.asarray(xy, dtype=np.float64)
    if xy.ndim == 2:
        xy = np.tile(xy[None], (len(targets), 1, 1))
    Q, N, _ = xy.shape
    stride = stride or N
    x = torch.zeros(Q, stride, dtype=torch.float64, device="cuda")
    y = torch.zeros(Q, stride, dtype=torch.float64, device="cuda")
    x[:, :N] = torch.as_tensor(xy[:, :, 0]).cuda()
    y[:, :N] = torch.as_tensor(xy[:, :, 1]).cuda()
    n = torch.full((Q,), N, dtype=torch.int32, device="cuda") if n_nodes is None else \
        torch.as_tensor(n_nodes, dtype=torch.int32).cuda()
    kw = {}
    if f32:
        kw = dict(x32=x.float(), y32=y.float(), coord_bound=22.0)
    idx, dist, tie = ops.nearest(x, y, n, targets, **kw)
    return np_(idx), np_(dist), np_(tie)


@pytest.mark.parametrize("f32", (False, True))
def test_nearest_golden(f32):
    g = golden("nearest")
    for n in (1, 7, 100, 1000, 5000):
        idx, dist, tie = _run(g[f"n{n}_xy"], g[f"n{n}_targets"], f32)
        assert np.array_equal(idx, g[f"n{n}_idx"])
        assert np.array_equal(dist, g[f"n{n}_dist"])      # sqrt(fma(dy,dy,dx*dx)): bit exact
    idx, _, _ = _run(g["tie_xy"], g["tie_targets"], f32)
    assert np.array_equal(idx, g["tie_idx"])              # duplicates: first index wins


@pytest.mark.parametrize("f32", (False, True))
def test_nearest_large_ragged_vs_oracle(f32):
    rng = np.random.default_rng(3)
    Q, N = 24, 60000
    xy = rng.uniform(-20, 20, (Q, N, 2))
    n_nodes = rng.integers(1, N, Q)
    n_nodes[0], n_nodes[1], n_nodes[2] = 1, N, 4
    t = rng.uniform(-20, 20, (Q, 2))
    idx, dist, tie = _run(xy, t, f32, n_nodes=n_nodes, stride=N + 4)
    oi, od = orc.nearest_batch(np.ascontiguousarray(xy[:, :, 0]), np.ascontiguousarray(xy[:, :, 1]), n_nodes, t)
    assert np.array_equal(idx, oi)
    assert np.array_equal(dist, od)


def test_nearest_clustered_candidates():
    """Many nodes inside the float32 guard band (overflow -> exact rescan path)."""
    rng = np.random.default_rng(4)
    N = 8192
    t = np.array([[1.0, 2.0]])
    ang = rng.uniform(0, 2 * np.pi, N)
    r = 3.0 + rng.uniform(0, 1e-9, N)            # a ring: every node is a near-tie in float32
    xy = np.stack([1.0 + r * np.cos(ang), 2.0 + r * np.sin(ang)], 1)
    idx, dist, tie = _run(xy, t, True)
    oi, od = orc.nearest(xy[:, 0], xy[:, 1], t[0])
    assert idx[0] == oi and dist[0] == od
