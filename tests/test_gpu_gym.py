"""DifferentialDriveGym / VectorDifferentialDriveGym (crlk_gym_step) vs episodes of the
unmodified reference and vs the oracle on a dense synthetic map."""
import numpy as np
import pytest

from conftest import golden
from gpu_util import make_envs, np_
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _maps():
    fx = golden("fixtures")
    return [fx[f"map_obs_list_{i}"] for i in range(8)]


@pytest.mark.parametrize("ep", range(5))
def test_gym_episode_golden(ep):
    """reset() with the reference's seed reproduces its start / goal / obstacle set, then every
    step's state, observation, reward, done flag and info entries."""
    from crl_kino_b200.env import DifferentialDriveEnv, DifferentialDriveGym
    from crl_kino_b200.utils.rigid import RectObs
    g = golden("gym")
    sets = [[RectObs(r.reshape(2, 2)) for r in m] for m in _maps()]
    gym_env = DifferentialDriveGym(DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi), obs_list_list=sets)
    gym_env.set_curriculum(ori=bool(g[f"ep{ep}_ori"]))
    np.random.seed(int(g[f"ep{ep}_seed"]))
    obs0 = gym_env.reset()
    assert np.array_equal(gym_env.robot_env.rects, golden("fixtures")["map_" + str(g[f"ep{ep}_map"])])
    np.testing.assert_allclose(gym_env.state, g[f"ep{ep}_start"], rtol=1e-12, atol=1e-14)
    assert np.array_equal(gym_env.goal, g[f"ep{ep}_goal"])
    assert np.array_equal(obs0[4:], g[f"ep{ep}_obs0"][4:])
    np.testing.assert_allclose(gym_env.a2v(g[f"ep{ep}_actions"][0]), g[f"ep{ep}_a2v"], rtol=1e-15)
    np.testing.assert_allclose(gym_env.v2a(g[f"ep{ep}_actions"][0]), g[f"ep{ep}_v2a"], rtol=1e-15)
    for t, a in enumerate(g[f"ep{ep}_actions"]):
        obs, rew, done, info = gym_env.step(a)
        np.testing.assert_allclose(gym_env.state, g[f"ep{ep}_states"][t], rtol=1e-9, atol=1e-11)
        assert np.array_equal(obs[4:], g[f"ep{ep}_obs"][t][4:]), (ep, t)
        np.testing.assert_allclose(obs[:4], g[f"ep{ep}_obs"][t][:4], rtol=1e-9, atol=1e-11)
        ref = g[f"ep{ep}_info"][t]
        assert list(info) == list(orc.GYM_INFO_KEYS)
        assert info["goal"] == bool(ref[0]) and info["collision"] == bool(ref[3])
        np.testing.assert_allclose([float(info[k]) for k in info], ref, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(rew, g[f"ep{ep}_reward"][t], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(gym_env._reward(info), rew, rtol=1e-12, atol=1e-12)
        assert done == bool(g[f"ep{ep}_done"][t])


def test_vector_gym_vs_oracle_dense_map():
    """512 environments per call on a 3,004-obstacle map: info, reward, done and observation of
    every environment against the oracle's scalar restatement."""
    from crl_kino_b200.env.differential_gym import VectorDifferentialDriveGym
    from crl_kino_b200.workloads import dense_map, free_states
    rects = dense_map(n_cells=3000, seed=4)
    env, o = make_envs(rects)
    B = 512
    rng = np.random.default_rng(9)
    states = free_states(lambda c: o.valid_state_batch(c)[1], rng, B, 0.0, moving=True)
    goals = free_states(lambda c: o.valid_state_batch(c)[1], rng, B, 0.5)
    goals[: B // 8, :2] = states[: B // 8, :2] + rng.uniform(-0.6, 0.6, (B // 8, 2))     # some at the goal
    venv = VectorDifferentialDriveGym(env, B)
    obs0 = np_(venv.set_states(states, goals))
    for step in range(3):
        actions = np.stack([rng.uniform(-0.1, 1.0, B), rng.uniform(-np.pi, np.pi, B)], 1)
        obs, rew, done, info = (np_(t) for t in venv.step(actions))
        new_states = np_(venv.states)
        n_near = 0
        for b in range(B):
            s, ob, r, inf = orc.gym_step(o, states[b], goals[b], actions[b])
            np.testing.assert_allclose(new_states[b], s, rtol=1e-9, atol=1e-11)
            if inf[3] != info[b, 3]:
                n_near += 1        # collision decision within 1e-6 of the threshold (bounded below)
                continue
            assert np.array_equal(ob[4:], obs[b, 4:]) or np.abs(ob[4:] - obs[b, 4:]).sum() <= 2
            np.testing.assert_allclose(info[b], inf, rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(rew[b], r, rtol=1e-9, atol=1e-8)
            assert bool(done[b]) == (bool(inf[0]) or bool(inf[3]))
        assert n_near <= 1
        states = new_states
    assert obs0.shape == (B, 733)


def test_vector_gym_reset_rule():
    """reset(): every environment gets a start and a goal with clearance > 0.5 and
    5 < ||start - goal|| < 8 (differential_gym.py:229-246); masked reset leaves the others alone."""
    from crl_kino_b200.env.differential_gym import VectorDifferentialDriveGym
    env, o = make_envs(_maps()[3])
    venv = VectorDifferentialDriveGym(env, 64, seed=5)
    obs = venv.reset()
    s, g = np_(venv.states), np_(venv.goals)
    d = np.linalg.norm(s[:, :2] - g[:, :2], axis=1)
    assert np.all((d > 5) & (d < 8))
    assert np.all(o.valid_state_batch(s)[1] > 0.5) and np.all(o.valid_state_batch(g)[1] > 0.5)
    assert tuple(obs.shape) == (64, 733)
    mask = np.zeros(64, dtype=bool)
    mask[::4] = True
    venv.reset(mask)
    s2 = np_(venv.states)
    assert np.array_equal(s2[~mask], s[~mask]) and not np.array_equal(s2[mask], s[mask])
