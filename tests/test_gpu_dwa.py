"""crlk_dwa_control vs the reference goldens and the oracle."""
import numpy as np
import pytest

from conftest import golden
from gpu_util import fixture_map, make_envs, np_
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

CONFS = {"rrt": dict(dt=0.2, v_res=0.02), "default": dict(),
         "dense": dict(dt=0.2, v_res=0.4 / 63, w_res=0.4 * np.pi / 63)}


def _dwa(env, **kw):
    from crl_kino_b200.policy.dwa import DWA
    d = DWA(env)
    d.set_dwa(**kw)
    return d


@pytest.mark.parametrize("name", list(CONFS))
def test_dwa_golden(name):
    g = golden("dwa")
    env, _ = make_envs(fixture_map("test_env1"))
    d = _dwa(env, **CONFS[name])
    r = d.control_batch(g[name + "_states"], g[name + "_goals"], want_traj=True)
    tie = np_(r["near_tie"]).astype(bool)
    assert np.array_equal(np_(r["n_rollouts"]), g[name + "_R"])
    idx = np_(r["opt_idx"])
    bad = (idx != g[name + "_idx"]) & ~tie
    assert not bad.any(), np.where(bad)
    same = idx == g[name + "_idx"]
    assert np.array_equal(np_(r["opt_v"])[same], g[name + "_optv"][same])      # grid values: bit exact
    np.testing.assert_allclose(np_(r["opt_cost"])[same], g[name + "_cmin"][same], rtol=1e-10)
    np.testing.assert_allclose(np_(r["opt_traj"])[same], g[name + "_traj"][same], rtol=1e-12, atol=1e-13)
    v, traj = d.control(g[name + "_states"][0], g[name + "_goals"][0])         # reference signature
    assert np.array_equal(v, g[name + "_optv"][0])
    assert traj.shape == g[name + "_traj"][0].shape


@pytest.mark.parametrize("mname", ("test_env1", "test_env2"))
def test_dwa_clearance_golden(mname):
    g = golden("dwa_clearance")
    env, _ = make_envs(fixture_map(mname))
    d = _dwa(env, dt=0.2, v_res=0.1, w_res=0.4, use_clearance=True)
    r = d.control_batch(g[mname + "_states"], g[mname + "_goals"])
    tie = np_(r["near_tie"]).astype(bool)
    assert np.array_equal(np_(r["n_rollouts"]), g[mname + "_R"])
    bad = (np_(r["opt_idx"]) != g[mname + "_idx"]) & ~tie
    assert not bad.any()
    np.testing.assert_allclose(np_(r["opt_cost"]), g[mname + "_cmin"], rtol=1e-8)


@pytest.mark.parametrize("name", ("rrt", "dense"))
def test_dwa_vs_oracle_random(name):
    from crl_kino_b200.workloads import random_states
    env, o = make_envs(fixture_map("test_env2"))
    d = _dwa(env, **CONFS[name])
    od = orc.Dwa(o, **CONFS[name])
    rng = np.random.default_rng(21)
    n = 400 if name == "rrt" else 96
    s = random_states(rng, n, moving=True)
    s[: n // 4, 3:] = 0
    goals = random_states(rng, n)
    r = d.control_batch(s, goals)
    ref = od.control_batch(s, goals)
    tie = np_(r["near_tie"]).astype(bool)
    assert np.array_equal(np_(r["n_rollouts"]), ref["R"])
    bad = (np_(r["opt_idx"]) != ref["opt_idx"]) & ~tie
    assert not bad.any(), (np.where(bad), tie.sum())
    assert tie.sum() <= max(2, n // 50)
    same = np_(r["opt_idx"]) == ref["opt_idx"]
    assert np.array_equal(np_(r["opt_v"])[same], ref["opt_v"][same])
    np.testing.assert_allclose(np_(r["opt_cost"])[same], ref["opt_cost"][same], rtol=1e-10)


def test_dwa_clearance_vs_oracle_dense_map():
    """Per-rollout clearance scoring (use_clearance=True) on a 10,004-obstacle map."""
    from crl_kino_b200.workloads import dense_map, free_states
    rects = dense_map()
    env, o = make_envs(rects)
    kw = dict(dt=0.2, v_res=0.05, w_res=0.25, use_clearance=True)
    d = _dwa(env, **kw)
    od = orc.Dwa(o, **kw)
    rng = np.random.default_rng(4)
    s = free_states(lambda c: o.valid_state_batch(c)[1], rng, 12, 0.3)
    goals = free_states(lambda c: o.valid_state_batch(c)[1], rng, 12, 0.3)
    r = d.control_batch(s, goals)
    ref = od.control_batch(s, goals)
    tie = np_(r["near_tie"]).astype(bool)
    bad = (np_(r["opt_idx"]) != ref["opt_idx"]) & ~tie
    assert not bad.any()
    np.testing.assert_allclose(np_(r["opt_cost"])[~tie], ref["opt_cost"][~tie], rtol=1e-8)


def test_rollout_atan2_accuracy():
    """The table-driven atan2 of the rollout loop (crlk_atan2) against numpy's arctan2 (the
    reference's, dwa.py:58) in units in the last place, over the argument ranges the planner
    produces plus axis / diagonal / tiny / zero cases."""
    import ctypes as C
    import torch
    from crl_kino_b200._lib import check, lib
    rng = np.random.default_rng(0)
    n = 400000
    y = rng.uniform(-45, 45, n)
    x = rng.uniform(-45, 45, n)
    y[:50000] = rng.uniform(-1e-4, 1e-4, 50000)                 # nearly along the x axis
    x[50000:100000] = rng.uniform(-1e-7, 1e-7, 50000)           # nearly along the y axis
    k = rng.integers(0, 17, 50000)
    y[100000:150000] = x[100000:150000] * (k / 16.0) * rng.choice([-1, 1], 50000)   # interval boundaries
    special = np.array([[0.0, 1.0], [0.0, -1.0], [1.0, 0.0], [-1.0, 0.0], [0.0, 0.0], [-0.0, 1.0], [1e-300, 1e-300],
                        [3.0, 3.0], [-3.0, 3.0], [3.0, -3.0], [-3.0, -3.0], [1e-310, 1.0], [5.0, 1e300]])
    y[:len(special)] = special[:, 0]
    x[:len(special)] = special[:, 1]
    dy, dx = torch.as_tensor(y).cuda(), torch.as_tensor(x).cuda()
    out = torch.empty_like(dy)
    check(lib().crlk_debug_atan2(C.c_void_p(dy.data_ptr()), C.c_void_p(dx.data_ptr()), n,
                                 C.c_void_p(out.data_ptr()), None))
    got = out.cpu().numpy()
    ref = np.arctan2(y, x)
    ulp = np.spacing(np.abs(ref))
    err = np.abs(got - ref) / np.where(ulp > 0, ulp, 1.0)
    assert np.all(np.signbit(got) == np.signbit(ref))
    assert err.max() <= 3.0, f"worst error {err.max()} ulp at y={y[err.argmax()]}, x={x[err.argmax()]}"
    assert np.mean(err <= 1.0) > 0.95
