"""Pins the oracle's gym restatement (oracle.gym_step / gym_info) against episodes of the
unmodified reference's DifferentialDriveGym (tests/golden/gym.npz).  CPU only."""
import numpy as np

from conftest import golden
from oracle import oracle as orc


def test_gym_episodes_match_reference():
    g = golden("gym")
    fx = golden("fixtures")
    for ep in range(int(g["n_episodes"])):
        env = orc.Env()
        env.set_obs(fx["map_" + str(g[f"ep{ep}_map"])])
        s, goal = g[f"ep{ep}_start"].copy(), g[f"ep{ep}_goal"]
        o0 = env.local_obs(s, goal)
        assert np.array_equal(o0[4:], g[f"ep{ep}_obs0"][4:])
        np.testing.assert_allclose(o0[:4], g[f"ep{ep}_obs0"][:4], rtol=1e-12, atol=1e-13)
        for t, a in enumerate(g[f"ep{ep}_actions"]):
            s, obs, rew, info = orc.gym_step(env, s, goal, a)
            assert np.array_equal(s, g[f"ep{ep}_states"][t]), (ep, t)
            assert np.array_equal(obs[4:], g[f"ep{ep}_obs"][t][4:]), (ep, t)      # occupancy map: exact
            np.testing.assert_allclose(obs[:4], g[f"ep{ep}_obs"][t][:4], rtol=1e-12, atol=1e-13)
            ref = g[f"ep{ep}_info"][t]
            assert info[0] == ref[0] and info[3] == ref[3], (ep, t)               # goal / collision flags
            np.testing.assert_allclose(info, ref, rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(rew, g[f"ep{ep}_reward"][t], rtol=1e-10, atol=1e-10)
            done = bool(info[0]) or bool(info[3])
            assert done == bool(g[f"ep{ep}_done"][t])
