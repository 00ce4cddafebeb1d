"""Properties that do not need the oracle, checked at BASELINE.json's full sizes: 100,000-node trees
(config 5) for the nearest-node scan, and the full 1024-query / 64 x 64-rollout batch (config 4) for
the fused extension."""
import numpy as np
import pytest
import torch

from gpu_util import make_envs, np_

pytestmark = pytest.mark.gpu


def test_nearest_100k_node_trees_properties():
    from crl_kino_b200 import ops
    Q, N = 256, 100000
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand(Q, N, generator=g, device="cuda", dtype=torch.float64) * 40 - 20
    y = torch.rand(Q, N, generator=g, device="cuda", dtype=torch.float64) * 40 - 20
    n = torch.full((Q,), N, dtype=torch.int32, device="cuda")
    x32, y32 = x.float(), y.float()
    # (1) the nearest node of a node's own position is that node (first occurrence), distance exactly 0
    pick = torch.randint(0, N, (Q,), generator=g, device="cuda")
    ar = torch.arange(Q, device="cuda")
    t = torch.stack([x[ar, pick], y[ar, pick]], 1)
    idx, dist, _ = ops.nearest(x, y, n, t, x32=x32, y32=y32, coord_bound=22.0)
    assert torch.equal(dist, torch.zeros_like(dist))
    assert torch.equal(x[ar, idx.long()], x[ar, pick]) and torch.equal(y[ar, idx.long()], y[ar, pick])
    assert bool((idx.long() <= pick).all())
    # (2) float32-screened scan == exact float64 scan, bit for bit, on random targets
    t = torch.rand(Q, 2, generator=g, device="cuda", dtype=torch.float64) * 40 - 20
    i32, d32, _ = ops.nearest(x, y, n, t, x32=x32, y32=y32, coord_bound=22.0)
    i64, d64, _ = ops.nearest(x, y, n, t)
    assert torch.equal(i32, i64) and torch.equal(d32, d64)
    # (3) the reported distance is the distance of the reported node and no node is closer
    dx = x - t[:, :1]
    dy = y - t[:, 1:]
    dd = torch.sqrt(torch.addcmul(dx * dx, dy, dy))
    assert torch.allclose(dd[ar, i64.long()], d64, rtol=1e-15, atol=0)
    assert bool((dd.min(dim=1).values >= d64 * (1 - 1e-15)).all())
    # (4) moving a tree's prefix length only ever picks nodes inside the prefix
    n2 = torch.randint(1, N, (Q,), generator=g, device="cuda").to(torch.int32)
    i2, _, _ = ops.nearest(x, y, n2, t, x32=x32, y32=y32, coord_bound=22.0)
    assert bool((i2 < n2).all())


def test_extend_full_cfg4_batch_properties():
    """1024 extensions on the 10,004-obstacle map with the nominal 64 x 64 rollout grid."""
    from crl_kino_b200 import ops
    from crl_kino_b200.policy.dwa import DWA
    from crl_kino_b200.workloads import DENSE_DWA, dense_map, free_states, random_states
    env, _ = make_envs(dense_map())
    rng = np.random.default_rng(12)
    clr = lambda c: np_(env.valid_state_batch(c, want_clearance=True)[1])
    Q = 1024
    froms = free_states(clr, rng, Q, 0.5)
    froms[:, 3:] = 0
    targets = random_states(rng, Q)
    d = DWA(env)
    d.set_dwa(**DENSE_DWA)
    xp = ops.extend_params()
    stats = torch.zeros(8, dtype=torch.int64, device="cuda")
    out = ops.extend_dwa(env.handle, d.params(), xp, froms, targets, want_ctrl=True, stats=stats)
    ns, status, nn = np_(out["n_steps"]), np_(out["status"]), np_(out["n_nodes"])
    path, node_step = np_(out["path"]), np_(out["node_step"])
    assert int(np_(stats)[0]) == Q and int(np_(stats)[1]) == int(ns.sum())
    assert ((ns >= 1) & (ns <= 50)).all()
    # every executed state but a colliding last one is valid; a "reached" path ends within 1 m
    for q in range(0, Q, 7):
        rows = path[q, :ns[q]]
        valid = np_(env.valid_state_batch(rows)[0]).astype(bool)
        assert valid[:-1].all()
        if status[q] == 1:                    # collision: the last state is the colliding one
            assert not valid[-1] or np_(out["flags"])[q] & 1
        if status[q] == 2:
            assert np.hypot(*(rows[-1, :2] - targets[q, :2])) <= 1.0 and node_step[q, nn[q] - 1] == ns[q]
        if status[q] == 0:
            assert ns[q] == 50 and nn[q] == 2 and list(node_step[q, :2]) == [25, 50]
        # consecutive states are one motion_velocity step apart: |dxy| <= v_max * dt
        seg = np.diff(np.vstack([froms[q][None], rows])[:, :2], axis=0)
        assert (np.hypot(seg[:, 0], seg[:, 1]) <= 0.2 + 1e-12).all()
    # determinism: the same launch twice gives the same bits (dynamic work distribution is invisible)
    out2 = ops.extend_dwa(env.handle, d.params(), xp, froms, targets, want_ctrl=True)
    for k in ("n_steps", "status", "n_nodes", "ctrl_idx"):
        assert torch.equal(out[k], out2[k]), k
    for q in range(0, Q, 13):
        assert np.array_equal(np_(out2["path"])[q, :ns[q]], path[q, :ns[q]])
    # re-running a single extension alone (latency launch shape, 768 threads) gives the same bits
    q = int(np.argmax(ns))
    one = ops.extend_dwa(env.handle, d.params(), xp, froms[q:q + 1], targets[q:q + 1], want_ctrl=True)
    assert int(np_(one["n_steps"])[0]) == ns[q]
    assert np.array_equal(np_(one["path"])[0, :ns[q]], path[q, :ns[q]])
    assert np.array_equal(np_(one["ctrl_idx"])[0, :ns[q]], np_(out["ctrl_idx"])[q, :ns[q]])
