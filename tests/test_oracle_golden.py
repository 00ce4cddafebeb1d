"""Pins the CPU oracle (oracle/) against vectors generated from the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden
from oracle import oracle as orc

MAPS = ("test_env1", "test_env2", "training_env")


def env_for(fixtures, mname):
    e = orc.Env()
    e.set_obs(fixtures["map_" + mname])
    return e


def test_fixture_constants(fixtures):
    assert np.array_equal(fixtures["robot_rect"], orc.ROBOT_RECT)
    assert np.array_equal(fixtures["action_low"], orc.ACTION_LOW)
    assert np.array_equal(fixtures["action_high"], orc.ACTION_HIGH)
    assert np.array_equal(fixtures["sample_positions"], orc.sample_positions())
    assert np.array_equal(fixtures["state_bounds"], orc.Env().get_bounds())


def test_arange_pairwise_normalize():
    g = golden("numerics")
    for s, sp, st, n, samp in zip(g["ar_start"], g["ar_span"], g["ar_step"], g["ar_len"], g["ar_samp"]):
        a = orc.arange(s, s + sp + st, st)
        assert len(a) == n
        if n >= 5:
            assert np.array_equal(np.concatenate([a[:3], a[-2:]]), samp)
    off = 0
    for n, val in zip(g["sum_n"], g["sum_val"]):
        x = g["sum_in"][off:off + n]
        off += n
        tab = np.zeros((n, 3))
        tab[:, 1] = x
        got = orc.lib().orc_pairwise_sum(tab[:, 1:].ctypes.data, int(n), 3)
        assert got == val, (n, got, val)
    out = np.array([orc.lib().orc_normalize_angle(float(a)) for a in g["ang_in"]])
    assert np.array_equal(out, g["ang_out"])


def test_motion_bit_exact():
    g = golden("motion")
    e = orc.Env()
    for i in range(len(g["states"])):
        s, c = g["states"][i], g["cmd"][i]
        assert np.array_equal(e.motion_velocity(s, c, 0.2), g["mv"][i])
        assert np.array_equal(e.calc_trajectory(s, c, 0.2, 1.0), g["traj"][i])
        assert np.array_equal(e.motion(s, g["acc"][i], 0.5), g["motion"][i])
        assert np.array_equal(e.dynamic_window(s, 0.2), g["dw"][i])
    for i in range(len(g["traj_dt01"])):
        assert np.array_equal(e.calc_trajectory(g["states"][i], g["cmd"][i], 0.1, 1.0), g["traj_dt01"][i])


def test_dis2obs_pairs():
    g = golden("geometry")
    e = orc.Env()
    got = np.array([e.dis2obs(p, r) for p, r in zip(g["pair_pose"], g["pair_rect"])])
    ref = g["pair_dis"]
    near = np.abs(np.abs(ref) - 0.05) < 1e-6
    col_ref, col_got = ref == -1, got == -1
    assert np.array_equal(col_ref[~near], col_got[~near])
    ok = ~col_ref & ~col_got
    assert ok.sum() > 500
    np.testing.assert_allclose(got[ok], ref[ok], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("mname", MAPS + ("obs_list_3",))
def test_valid_and_clearance(fixtures, mname):
    g = golden("geometry")
    e = env_for(fixtures, mname)
    s = g[mname + "_states"]
    valid, clr = e.valid_state_batch(s, threads=2)
    assert np.array_equal(valid, g[mname + "_valid"])
    np.testing.assert_allclose(clr, g[mname + "_clear"], rtol=1e-9, atol=1e-12)
    assert np.array_equal(e.valid_point_check(g[mname + "_pts"]), g[mname + "_pts_valid"])
    assert e.valid_state_check(s[0]) == bool(g[mname + "_valid"][0])


@pytest.mark.parametrize("mname", ("test_env1", "test_env2"))
def test_local_obs(fixtures, mname):
    g = golden("obs")
    e = env_for(fixtures, mname)
    got = e.local_obs_batch(g[mname + "_states"], g[mname + "_goals"], threads=2)
    ref = g[mname + "_obs"]
    assert np.array_equal(got[:, 4:], ref[:, 4:])
    np.testing.assert_allclose(got[:, :4], ref[:, :4], rtol=1e-10, atol=1e-12)
    assert np.array_equal(e.local_obs(g[mname + "_states"][3], g[mname + "_goals"][3]), got[3])


def test_nearest():
    g = golden("nearest")
    for n in (1, 7, 100, 1000, 5000):
        xy = g[f"n{n}_xy"]
        for t, i, d in zip(g[f"n{n}_targets"], g[f"n{n}_idx"], g[f"n{n}_dist"]):
            gi, gd = orc.nearest(xy[:, 0], xy[:, 1], t)
            assert gi == i and gd == d
    xy = g["tie_xy"]
    for t, i in zip(g["tie_targets"], g["tie_idx"]):
        assert orc.nearest(xy[:, 0], xy[:, 1], t)[0] == i
    xs = np.tile(xy[:, 0], (3, 1)).copy()
    ys = np.tile(xy[:, 1], (3, 1)).copy()
    idx, _ = orc.nearest_batch(xs, ys, [100, 100, 50], g["tie_targets"][:3], threads=2)
    assert idx[0] == g["tie_idx"][0]


def _seeded_actor(seed):
    import torch
    torch.manual_seed(seed)
    dims = [733, 1024, 512, 512, 512, 2]
    Ws, bs = [], []
    for i in range(5):
        lin = torch.nn.Linear(dims[i], dims[i + 1])
        Ws.append(lin.weight.detach().numpy())
        bs.append(lin.bias.detach().numpy())
    return Ws, bs


def test_actor_forward():
    g = golden("policy")
    Ws, bs = _seeded_actor(int(g["seed"]))
    fp = np.array([float(np.sum(W.astype(np.float64))) for W in Ws] +
                  [float(np.sum(b.astype(np.float64))) for b in bs])
    np.testing.assert_allclose(fp, g["weight_fingerprint"], rtol=0, atol=0)
    a0 = orc.actor_forward(Ws, bs, g["obs"])
    np.testing.assert_allclose(a0, g["act_eps0"], rtol=2e-5, atol=2e-6)
    a1 = orc.actor_forward(Ws, bs, g["obs"], g["noise"], 0.05)
    np.testing.assert_allclose(a1, g["act_eps005"], rtol=2e-5, atol=2e-6)


def test_estimator_forward():
    g = golden("estimator")
    w = golden("estimator_weights")
    est = {k[4:]: w[k] for k in w.files if k.startswith("est.")}
    cls = {k[4:]: w[k] for k in w.files if k.startswith("cls.")}
    sp, ext = orc.estimator_inputs(g["obs"])
    np.testing.assert_allclose(orc.cnn_forward(est, sp, ext, False)[:, 0], g["est"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(orc.cnn_forward(cls, sp, ext, True)[:, 0], g["cls"], rtol=1e-5, atol=1e-6)


def test_estimator_choose_parent(fixtures):
    g = golden("estimator")
    w = golden("estimator_weights")
    est = {k[4:]: w[k] for k in w.files if k.startswith("est.")}
    cls = {k[4:]: w[k] for k in w.files if k.startswith("cls.")}
    e = env_for(fixtures, "test_env2")
    p = orc.PlannerPort(e, steering="rl", estimator=est, classifier=cls)
    for k in range(3):
        nodes = [orc.Node(s) for s in g[f"cp{k}_nodes"]]
        for t, i, c in zip(g[f"cp{k}_targets"], g[f"cp{k}_idx"], g[f"cp{k}_cost"]):
            gi, gc = p.choose_parent(nodes, orc.Node(t))
            assert gi == i
            assert gc == pytest.approx(c, rel=1e-5)


def test_circle_robot_golden(fixtures):
    """CircleRobot footprint (rigid.py:331-347)."""
    g = golden("circle")
    radius = float(g["radius"])
    e = orc.Env()
    for pose, rect, ref in zip(g["pair_pose"], g["pair_rect"], g["pair_dis"]):
        got = e.dis2obs_circle(radius, pose, rect)
        assert (got == -1.0) == (ref == -1.0)
        if ref != -1.0:
            assert abs(got - ref) <= 1e-9 * max(1.0, abs(ref))
    for mname in ("test_env1", "test_env2"):
        env = orc.Env(robot_radius=radius)
        env.set_obs(fixtures["map_" + mname])
        valid, clr = env.valid_state_batch(g[f"{mname}_states"])
        assert np.array_equal(valid, g[f"{mname}_valid"])
        np.testing.assert_allclose(clr, g[f"{mname}_clear"], rtol=1e-9, atol=1e-12)
