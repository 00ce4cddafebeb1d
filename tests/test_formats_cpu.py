"""The reference's on-disk formats through crl_kino_b200.utils.io: files
pickled by the unmodified reference (bytes kept in tests/golden/formats.npz) are read without
the reference installed, and files written here carry the reference's class paths (the
generating script checked that the reference's own pickle.load accepts them).  CPU only."""
import pickle
import sys
import types

import numpy as np

from conftest import golden
from crl_kino_b200.utils import io as cio


def test_read_reference_pickles():
    g = golden("formats")
    assert "crl_kino" not in sys.modules                      # read without the reference importable
    obs = cio.load_obstacles(g["obs_pickle"].tobytes())
    assert type(obs[0]).__name__ == "RectObs" and obs[0].rect.dtype == np.float32
    assert np.array_equal(cio.load_obstacles_flat(g["obs_pickle"].tobytes()), g["obs_rects"])
    lol = cio.load_obstacles_flat(g["obs_lol_pickle"].tobytes())
    assert np.array_equal(lol[0], g["obs_lol_rects0"]) and np.array_equal(lol[1], g["obs_lol_rects1"])
    assert np.array_equal(cio.load_positions(g["positions_pickle"].tobytes()), g["positions"])
    nodes = cio.load_tree(g["tree_pickle"].tobytes())
    states, parents = cio.tree_arrays(nodes)
    assert np.array_equal(states, g["tree_states"]) and np.array_equal(parents, g["tree_parents"])
    assert np.array_equal([len(n.path) for n in nodes], g["tree_plen"])
    assert np.array_equal(cio.load_path(g["path_pickle"].tobytes()), g["path"])
    res = cio.load_results(g["results_pickle"].tobytes())
    assert set(res) == {"runtime", "path_len"} and res["runtime"][1] == -1


def test_written_files_carry_reference_class_paths(tmp_path):
    g = golden("formats")
    blob = cio.save_obstacles(g["obs_rects"], str(tmp_path / "map.pkl"))
    assert b"crl_kino.utils.rigid" in blob and b"RectObs" in blob and b"crl_kino_b200" not in blob
    assert "crl_kino" not in sys.modules                      # the stand-in modules are gone again
    assert np.array_equal(cio.load_obstacles_flat(str(tmp_path / "map.pkl")), g["obs_rects"])
    nodes = cio.load_tree(g["tree_pickle"].tobytes())
    tblob = cio.save_tree(nodes)
    assert b"crl_kino.planner.rrt" in tblob and b"RRT.Node" in tblob and b"crl_kino_b200" not in tblob
    # a stand-in for the reference's modules (plain classes, as pickle needs) loads both files
    mods = {n: types.ModuleType(n) for n in ("crl_kino", "crl_kino.utils", "crl_kino.utils.rigid",
                                             "crl_kino.planner", "crl_kino.planner.rrt")}
    mods["crl_kino.utils.rigid"].RectObs = type("RectObs", (), {"__module__": "crl_kino.utils.rigid"})
    node = type("Node", (), {"__module__": "crl_kino.planner.rrt", "__qualname__": "RRT.Node"})
    mods["crl_kino.planner.rrt"].RRT = type("RRT", (), {"Node": node})
    sys.modules.update(mods)
    try:
        back = pickle.loads(blob)
        assert np.array_equal(np.array([o.rect.reshape(4) for o in back]), g["obs_rects"]) and back[0].color == "black"
        tb = pickle.loads(tblob)
        assert tb[1].parent is tb[int(g["tree_parents"][1])] and np.array_equal(tb[-1].state, g["tree_states"][-1])
    finally:
        for n in mods:
            sys.modules.pop(n, None)


def test_planning_results_dict(tmp_path):
    """examples/benchmark.py:38-43, :69-75, :86-88."""
    data = cio.planning_results([True, False, True], [1.5, 9.0, 2.5], [63, 10, 151])
    assert np.array_equal(data["runtime"], [1.5, -1, 2.5])
    np.testing.assert_allclose(data["path_len"], [0.2 * 62, 0.2 * 150])
    cio.save_results(data, str(tmp_path / "planning_results_env1"))
    back = pickle.load(open(tmp_path / "planning_results_env1.pkl", "rb"))
    assert np.array_equal(back["runtime"], data["runtime"]) and np.array_equal(back["path_len"], data["path_len"])
    assert cio.benchmark_pairs(np.zeros((4, 5)))[:4] == [(0, 1), (0, 2), (0, 3), (1, 0)]


def test_gym_action_space_maps_match_reference():
    """DifferentialDriveGym.a2v / v2a (differential_gym.py:110-128) are host arithmetic: against the
    values the unmodified reference produced (tests/golden/gym.npz)."""
    from crl_kino_b200.env import DifferentialDriveEnv, DifferentialDriveGym
    g = golden("gym")
    gym_env = DifferentialDriveGym(DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi))
    for ep in range(int(g["n_episodes"])):
        a = g[f"ep{ep}_actions"][0]
        assert np.array_equal(gym_env.a2v(a), g[f"ep{ep}_a2v"])
        assert np.array_equal(gym_env.v2a(a), g[f"ep{ep}_v2a"])
    assert gym_env.observation_space.shape == (733,) and gym_env.action_space.low.dtype == np.float32
    assert gym_env.curriculum == {"obs_num": 7, "ori": False}
    gym_env.set_curriculum(ori=True)
    assert gym_env.curriculum["ori"] is True
    info = dict(zip(("goal", "goal_dis", "heading", "collision", "clearance", "v", "w", "step"),
                    (True, 0.5, 0.1, False, 1.0, 0.7, 0.2, 1.0)))
    assert abs(gym_env._reward(info) - (50.0 - 0.75 - 0.05 + 1.0 + 0.7 - 0.1 - 2.5)) < 1e-12


def test_policy_state_dict_in_the_reference_layout():
    """load_policy's load_state_dict takes the reference's DDPGPolicy state dict (actor / actor_old /
    critic / critic_old entries; the actor is nn.Sequential `model` with Linear layers at even
    indices, policy/net.py:126-133) and keeps only the actor."""
    import torch
    from crl_kino_b200.env import DifferentialDriveEnv, DifferentialDriveGym
    from crl_kino_b200.policy.rl_policy import DDPGPolicy
    from crl_kino_b200.policy.net import Actor
    gym_env = DifferentialDriveGym(DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi))
    layer = [1024, 512, 512, 512]
    dims = [733] + layer + [2]
    rng = np.random.default_rng(0)
    sd = {}
    for prefix in ("actor", "actor_old"):
        for i in range(5):
            sd[f"{prefix}.model.{2 * i}.weight"] = torch.as_tensor(rng.standard_normal((dims[i + 1], dims[i])).astype(np.float32))
            sd[f"{prefix}.model.{2 * i}.bias"] = torch.as_tensor(rng.standard_normal(dims[i + 1]).astype(np.float32))
    sd["critic.model.0.weight"] = torch.zeros(4, 735)
    sd["critic_old.model.0.weight"] = torch.zeros(4, 735)
    actor = Actor(layer, gym_env.observation_space.shape, gym_env.action_space.shape,
                  [gym_env.action_space.low, gym_env.action_space.high], seed=1)
    policy = DDPGPolicy(actor, None, None, None, action_range=[gym_env.action_space.low, gym_env.action_space.high])
    policy.load_state_dict(sd)
    for i in range(5):
        assert np.array_equal(actor.weights[i], sd[f"actor.model.{2 * i}.weight"].numpy())
        assert np.array_equal(actor.biases[i], sd[f"actor.model.{2 * i}.bias"].numpy())
    assert actor.low.dtype == np.float32 and np.allclose(actor.low, [-0.1, -np.pi]) and np.allclose(actor.high, [1.0, np.pi])


def test_device_sample_stream_host_mirror_properties():
    """The host mirror of the planner kernel's sample generator (workloads.device_sample_stream):
    deterministic in (seed, query, draw), independent of the batch it is asked in, inside the state
    bounds, goal bias (5 + 1) / 101 (rrt.py:147)."""
    from crl_kino_b200.workloads import device_sample_stream
    sb = golden("fixtures")["state_bounds"]
    goals = np.tile(np.array([[10, 10, 0, 0, 0.0]]), (6, 1)) + np.arange(6)[:, None] * 0.1
    a = device_sample_stream(sb, goals, 3000, seed=7)
    b = device_sample_stream(sb, goals[:3], 1000, seed=7)
    assert a.shape == (6, 3000, 5) and np.array_equal(a[:3, :1000], b)
    assert not np.array_equal(a, device_sample_stream(sb, goals, 3000, seed=8))
    is_goal = np.all(a == goals[:, None, :], axis=2)
    assert 0.045 < is_goal.mean() < 0.075
    u = a[~is_goal]
    assert np.all(u >= sb[:, 0]) and np.all(u < sb[:, 1])
    assert abs(u[:, 0].mean()) < 0.5 and 10.5 < u[:, 0].std() < 12.5          # uniform on [-20, 20)
