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
    np.testing.assert_allclose(np_(r["opt_cost"])[same], g[name + "_cmin"][same], rtol=1e-13)
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
    np.testing.assert_allclose(np_(r["opt_cost"])[same], ref["opt_cost"][same], rtol=1e-13)


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


def test_rollout_atan2_matches_the_c_library():
    """crlk_atan2 (correctly rounded, csrc/crlk_crmath.h) against math.atan2 -- the C library call
    that np.arctan2 makes for the reference's Python scalars (dwa.py:58) -- bit for bit, over the
    argument ranges the planner produces plus axis / diagonal / tiny / zero cases.  The library itself
    misrounds 7.5e-4 of random arguments (tools/cr_check.cpp): those may differ, by one ulp."""
    import ctypes as C
    import math
    import torch
    from crl_kino_b200._lib import check, lib
    rng = np.random.default_rng(0)
    n = 400000
    y = rng.uniform(-45, 45, n)
    x = rng.uniform(-45, 45, n)
    y[:50000] = rng.uniform(-1e-4, 1e-4, 50000)                 # nearly along the x axis
    x[50000:100000] = rng.uniform(-1e-7, 1e-7, 50000)           # nearly along the y axis
    k = rng.integers(0, 65, 50000)
    y[100000:150000] = x[100000:150000] * (k / 64.0) * rng.choice([-1, 1], 50000)   # interval boundaries
    special = np.array([[0.0, 1.0], [0.0, -1.0], [1.0, 0.0], [-1.0, 0.0], [0.0, 0.0], [-0.0, 1.0], [1e-300, 1e-300],
                        [3.0, 3.0], [-3.0, 3.0], [3.0, -3.0], [-3.0, -3.0], [1e-310, 1.0], [5.0, 1e300],
                        [0.0, -0.0], [-0.0, -0.0], [1e-200, 1e200], [1e200, 1e-200]])
    y[:len(special)] = special[:, 0]
    x[:len(special)] = special[:, 1]
    dy, dx = torch.as_tensor(y).cuda(), torch.as_tensor(x).cuda()
    out = torch.empty_like(dy)
    check(lib().crlk_debug_atan2(C.c_void_p(dy.data_ptr()), C.c_void_p(dx.data_ptr()), n,
                                 C.c_void_p(out.data_ptr()), None))
    got = out.cpu().numpy()
    ref = np.array([math.atan2(a, b) for a, b in zip(y, x)])
    ulp = np.spacing(np.abs(ref))
    err = np.abs(got - ref) / np.where(ulp > 0, ulp, 1.0)
    assert np.all(np.signbit(got) == np.signbit(ref))
    assert err.max() <= 1.0, f"worst error {err.max()} ulp at y={y[err.argmax()]}, x={x[err.argmax()]}"
    assert np.mean(got == ref) > 0.998, np.mean(got == ref)


def test_heading_sincos_matches_the_c_library():
    """crlk_sincos against math.sin / math.cos (differential_env.py:98-99), bit for bit on normalised
    headings (the library misrounds ~6e-4 of them; one ulp there)."""
    import ctypes as C
    import math
    import torch
    from crl_kino_b200._lib import check, lib
    rng = np.random.default_rng(1)
    n = 300000
    th = rng.uniform(-np.pi, np.pi, n)
    th[:20000] = rng.uniform(-1e-5, 1e-5, 20000)
    th[20000:40000] = np.pi - rng.uniform(0, 1e-5, 20000)
    th[40000:60000] = np.pi / 2 + rng.uniform(-1e-5, 1e-5, 20000)
    th[60000:60202] = np.arange(202) / 64.0 * rng.choice([-1, 1], 202)
    th[60202:60208] = [0.0, -0.0, np.pi, -np.pi, np.pi / 2, 5.0]        # 5.0: outside the table, library path
    d = torch.as_tensor(th).cuda()
    sn, cs = torch.empty_like(d), torch.empty_like(d)
    p = lambda t: C.c_void_p(t.data_ptr())
    check(lib().crlk_debug_sincos(p(d), n, p(sn), p(cs), None))
    gs, gc = sn.cpu().numpy(), cs.cpu().numpy()
    rs = np.array([math.sin(a) for a in th])
    rc = np.array([math.cos(a) for a in th])
    for got, ref in ((gs, rs), (gc, rc)):
        ulp = np.spacing(np.abs(ref))
        err = np.abs(got - ref) / np.where(ulp > 0, ulp, 1.0)
        assert err.max() <= 2.0, err.max()
        assert np.mean(got == ref) > 0.998, np.mean(got == ref)


def test_cost_normalisation_quotient_is_correctly_rounded():
    """The cost columns are divided by their sums with one multiply and two fused corrections on the
    reciprocal (div_rn_y); it must return exactly what the division operator returns (dwa.py:78)."""
    import ctypes as C
    import torch
    from crl_kino_b200._lib import check, lib
    rng = np.random.default_rng(3)
    n = 1 << 21
    a = rng.uniform(0.0, np.pi, n)
    b = rng.uniform(1e-3, 2e4, n)
    a[:1000] = 0.0
    a[1000:2000] = rng.uniform(-1.0, 1.0, 1000)                       # speed costs can be negative
    b[2000:3000] = np.nextafter(2.0 ** rng.integers(-3, 14, 1000), 0)  # significands of all ones
    a[3000:4000] = b[3000:4000]
    a[4000:300000] = rng.uniform(0, 1, 296000) ** 8 * np.pi            # many small heading costs
    da, db = torch.as_tensor(a).cuda(), torch.as_tensor(b).cuda()
    fast, exact = torch.empty_like(da), torch.empty_like(da)
    p = lambda t: C.c_void_p(t.data_ptr())
    check(lib().crlk_debug_div(p(da), p(db), n, p(fast), p(exact), None))
    assert np.array_equal(np_(exact), a / b)
    assert np.array_equal(np_(fast), a / b)


def test_dwa_empty_window_is_reported():
    """v outside the limits by more than acc * dt: np.arange is empty and the reference raises
    (dwa.py:76 on the empty cost table); here opt_idx = -1 / n_rollouts = 0 / NaN, and an extension
    from such a state ends at once with CRLK_EXT_INVALID and flag bit 3 (no out-of-range reads)."""
    from crl_kino_b200 import ops
    env, _ = make_envs(fixture_map("test_env1"))
    d = _dwa(env, dt=0.2, v_res=0.02)
    s = np.array([[0.0, 0.0, 0.0, 2.0, 0.0], [0.0, 0.0, 0.0, 0.3, 0.0], [0.0, 0.0, 0.0, 0.0, 9.0]])
    g = np.array([[5.0, 5.0, 0, 0, 0]] * 3)
    r = d.control_batch(s, g, want_traj=True)
    assert np_(r["opt_idx"]).tolist()[0] == -1 and np_(r["opt_idx"])[2] == -1 and np_(r["opt_idx"])[1] >= 0
    assert np_(r["n_rollouts"]).tolist()[0] == 0 and np_(r["n_rollouts"])[1] > 0
    assert np.isnan(np_(r["opt_v"])[0]).all() and np.isnan(np_(r["opt_traj"])[0]).all()
    assert np.isfinite(np_(r["opt_v"])[1]).all()
    out = ops.extend_dwa(env.handle, d.params(), ops.extend_params(), s, g)
    assert np_(out["status"]).tolist()[0] == 3 and np_(out["n_steps"])[0] == 0 and np_(out["flags"])[0] & 8
    assert np_(out["status"])[2] == 3
    assert np_(out["n_steps"])[1] > 0 and not (np_(out["flags"])[1] & 8)
    from crl_kino_b200.planner.rrt import RRT
    p = RRT(env)
    with pytest.raises(ValueError):
        p.set_start_and_goal(np.array([-5, -15, 0, 2.0, 0.0]), np.array([10, 10, 0, 0, 0.0]))
