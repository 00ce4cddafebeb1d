"""CUDA env kernels (through the C ABI) vs the reference goldens and the oracle."""
import numpy as np
import pytest

from conftest import golden
from gpu_util import assert_flags, fixture_map, make_envs, np_

pytestmark = pytest.mark.gpu


def test_library_loaded():
    from crl_kino_b200._lib import lib
    assert lib().crlk_device_count() >= 1
    assert lib().crlk_abi_version() == 2


def test_motion_golden():
    g = golden("motion")
    env, o = make_envs(np.zeros((0, 4), dtype=np.float32))
    mv = np_(env.motion_velocity_batch(g["states"], g["cmd"], 0.2))
    # trajectories: 1e-4 relative is the contract; the kernels are far tighter
    np.testing.assert_allclose(mv, g["mv"], rtol=1e-12, atol=1e-13)
    exact = float(np.mean(mv == g["mv"]))
    assert exact > 0.9, exact
    from crl_kino_b200 import ops
    mo = np_(ops.motion_velocity(env.handle, g["states"], g["acc"], 0.5, accel=True))
    np.testing.assert_allclose(mo, g["motion"], rtol=1e-12, atol=1e-13)
    dw = np_(ops.dynamic_window(env.handle, g["states"], 0.2)).reshape(-1, 2, 2)
    assert np.array_equal(dw, g["dw"])
    # single-call API keeps the reference signatures
    s1 = env.motion_velocity(g["states"][0], g["cmd"][0], 0.2)
    np.testing.assert_allclose(s1, g["mv"][0], rtol=1e-12, atol=1e-13)
    assert np.array_equal(env.dynamic_window(g["states"][0], 0.2), g["dw"][0])
    from crl_kino_b200.policy.dwa import DWA
    d = DWA(env)
    d.set_dwa(dt=0.2, v_res=0.02)
    np.testing.assert_allclose(d.calc_trajectory(g["states"][3], g["cmd"][3]), g["traj"][3], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("mname", ("test_env1", "test_env2", "training_env", "obs_list_3"))
def test_valid_clearance_golden(mname):
    g = golden("geometry")
    env, o = make_envs(fixture_map(mname))
    s = g[mname + "_states"]
    valid, clr, near = env.valid_state_batch(s, want_clearance=True, want_near=True)
    assert_flags(np_(valid), g[mname + "_valid"], np_(near), "valid_state_check")
    ref = g[mname + "_clear"]
    ok = (ref > 0) & (np_(clr) > 0)
    np.testing.assert_allclose(np_(clr)[ok], ref[ok], rtol=1e-9)
    assert np.array_equal(np_(clr) < 0, ref < 0) or np_(near).any()
    # flag-only path must agree with the clearance path
    v2, _, _ = env.valid_state_batch(s)
    assert np.array_equal(np_(v2), np_(valid))
    pv = env.valid_point_check(g[mname + "_pts"])
    assert np.array_equal(pv, g[mname + "_pts_valid"])
    assert env.valid_state_check(s[0]) == bool(g[mname + "_valid"][0])
    assert env.get_clearance(s[1]) == pytest.approx(ref[1], rel=1e-9)


def test_dis2obs_pairs_golden():
    """Single-obstacle envs reproduce RectRobot.dis2obs incl. the obstacle-inside-robot quirk."""
    g = golden("geometry")
    poses, rects, ref = g["pair_pose"], g["pair_rect"], g["pair_dis"]
    keys = {}
    for i, r in enumerate(rects):
        keys.setdefault(r.tobytes(), []).append(i)
    n_near = 0
    for k, idx in keys.items():
        idx = np.array(idx)
        env, _ = make_envs(rects[idx[0]][None])
        st = np.zeros((len(idx), 5))
        st[:, :3] = poses[idx]
        valid, clr, near = env.valid_state_batch(st, want_clearance=True, want_near=True)
        clr, near = np_(clr), np_(near)
        n_near += assert_flags(clr < 0, ref[idx] < 0, near, "dis2obs collision")
        ok = (clr > 0) & (ref[idx] > 0)
        np.testing.assert_allclose(clr[ok], ref[idx][ok], rtol=1e-8, atol=1e-10)
    assert n_near <= 2


@pytest.mark.parametrize("mname", ("test_env1", "test_env2"))
def test_local_obs_golden(mname):
    g = golden("obs")
    env, o = make_envs(fixture_map(mname))
    from crl_kino_b200.env import DifferentialDriveGym
    gym = DifferentialDriveGym(env)
    o64, o32, near = gym.obs_batch(g[mname + "_states"], g[mname + "_goals"], want32=True, want_near=True)
    o64, o32, near = np_(o64), np_(o32), np_(near)
    ref = g[mname + "_obs"]
    rows = ~near.astype(bool)
    assert rows.sum() >= len(ref) - 4
    assert np.array_equal(o64[rows, 4:], ref[rows, 4:])
    np.testing.assert_allclose(o64[:, :4], ref[:, :4], rtol=1e-10, atol=1e-12)
    assert np.array_equal(o32[:, :733], o64.astype(np.float32))
    assert not o32[:, 733:].any()
    gym.state, gym.goal = g[mname + "_states"][5], g[mname + "_goals"][5]
    assert np.array_equal(gym._obs(), o64[5])
    assert np.array_equal(gym.sample_positions, golden("fixtures")["sample_positions"])


def test_dense_map_vs_oracle():
    """cfg4-sized map (10,004 obstacles): flags / clearance / observations vs the oracle."""
    from crl_kino_b200.workloads import dense_map, random_states
    rects = dense_map()
    env, o = make_envs(rects)
    rng = np.random.default_rng(5)
    s = random_states(rng, 1500, moving=True)
    valid, clr, near = env.valid_state_batch(s, want_clearance=True, want_near=True)
    ov, oc = o.valid_state_batch(s)
    assert_flags(np_(valid), ov, np_(near), "dense valid")
    ok = (oc > 0) & (np_(clr) > 0)
    assert ok.sum() > 300
    np.testing.assert_allclose(np_(clr)[ok], oc[ok], rtol=1e-8, atol=1e-10)
    v2, _, _ = env.valid_state_batch(s)
    assert np.array_equal(np_(v2), np_(valid))
    from crl_kino_b200 import ops
    g = random_states(rng, 64)
    o64, _, nb = ops.local_obs(env.handle, s[:64], g, want_near=True)
    ref = o.local_obs_batch(s[:64], g)
    rows = ~np_(nb).astype(bool)
    assert np.array_equal(np_(o64)[rows, 4:], ref[rows, 4:])
    np.testing.assert_allclose(np_(o64)[:, :4], ref[:, :4], rtol=1e-10, atol=1e-12)


def test_edge_cases():
    env, o = make_envs(np.zeros((0, 4), dtype=np.float32))     # no obstacles at all
    s = np.zeros((3, 5))
    valid, clr, _ = env.valid_state_batch(s, want_clearance=True)
    assert np_(valid).all() and np.all(np_(clr) == 100.0)
    assert env.valid_point_check(np.zeros((2, 2))).all()
    from crl_kino_b200 import ops
    v, c, n = ops.valid_state(env.handle, np.zeros((0, 5)))
    assert v.numel() == 0
    # obstacle strictly inside the footprint is NOT a collision
    env2, o2 = make_envs(np.array([[-0.1, -0.1, 0.1, 0.1]], dtype=np.float32))
    assert env2.valid_state_check(np.zeros(5)) == o2.valid_state_check(np.zeros(5)) == True  # noqa: E712
    assert env2.get_clearance(np.zeros(5)) == pytest.approx(o2.get_clearance(np.zeros(5)), rel=1e-9)
    # add_obs / set_obs refresh the device copy
    from crl_kino_b200.utils.rigid import RectObs
    env2.add_obs(RectObs(np.array([[0.5, -0.2], [3.0, 0.2]])))
    assert not env2.valid_state_check(np.zeros(5))
    with pytest.raises(AssertionError):
        env2.add_obs("not an obstacle")


@pytest.mark.parametrize("mname", ("test_env1", "test_env2"))
def test_circle_robot_golden(mname):
    """env.rigid_robot = CircleRobot(radius) (rigid.py:331-347): validity and clearance against
    the unmodified reference; near-threshold states are reported, not compared."""
    from crl_kino_b200.utils.rigid import CircleRobot
    g = golden("circle")
    env, _ = make_envs(fixture_map(mname))
    env.rigid_robot = CircleRobot(float(g["radius"]))
    valid, clr, near = env.valid_state_batch(g[f"{mname}_states"], want_clearance=True, want_near=True)
    assert_flags(np_(valid), g[f"{mname}_valid"], np_(near), "circle valid")
    ok = np_(near) == 0
    np.testing.assert_allclose(np_(clr)[ok], g[f"{mname}_clear"][ok], rtol=1e-9, atol=1e-12)
    assert env.valid_state_check(g[f"{mname}_states"][0]) == bool(g[f"{mname}_valid"][0])
    # one robot-obstacle pair at a time: a single-obstacle env per golden pair
    env1, _ = make_envs(g["pair_rect"][:1])
    env1.rigid_robot = CircleRobot(float(g["radius"]))
    for k in range(0, len(g["pair_dis"]), 17):
        env1.set_obs(g["pair_rect"][k:k + 1])
        s = np.zeros(5)
        s[:3] = g["pair_pose"][k]
        c = env1.get_clearance(s)
        ref = g["pair_dis"][k]
        if abs(abs(ref) - 0.05) > 1e-6:
            assert (c == -1.0) == (ref == -1.0)
            if ref != -1.0:
                assert abs(c - min(ref, 100.0)) <= 1e-9 * max(1.0, abs(ref))


@pytest.mark.parametrize("seed,cells", ((0, 10000), (5, 2500)))
def test_local_obs_raster_path_equals_exact_path(seed, cells):
    """Observations built through the occupancy raster (no near-boundary output requested) are
    bit-identical to the exact per-obstacle path (near-boundary output requested) and to the
    oracle's scan over all obstacles, also for poses hugging obstacle faces and raster lines."""
    from crl_kino_b200 import ops
    from crl_kino_b200.workloads import dense_map, random_states
    rects = dense_map(n_cells=cells, seed=seed)
    env, o = make_envs(rects)
    rng = np.random.default_rng(seed + 1)
    B = 3000
    states = random_states(rng, B, moving=True)
    goals = random_states(rng, B)
    # a third of the poses exactly on raster / obstacle lines, axis aligned: sample points land on cell edges
    k = B // 3
    states[:k, 0] = np.round(states[:k, 0] / 0.125) * 0.125
    states[:k, 1] = np.round(states[:k, 1] / 0.125) * 0.125
    states[:k, 2] = rng.choice([0.0, np.pi / 2, -np.pi / 2, np.pi], k)
    fast, f32, _ = ops.local_obs(env.handle, states, goals, want64=True, want32=True, want_near=False)
    exact, _, near = ops.local_obs(env.handle, states, goals, want64=True, want32=False, want_near=True)
    assert np.array_equal(np_(fast), np_(exact))
    assert np.array_equal(np_(f32)[:, :733], np_(fast).astype(np.float32))
    ref = o.local_obs_batch(states[:400], goals[:400])
    ok = np_(near)[:400] == 0
    assert np.array_equal(np_(fast)[:400][ok][:, 4:], ref[ok][:, 4:])
    assert 0.02 < np_(fast)[:, 4:].mean() < 0.98          # both free and occupied cells occur
