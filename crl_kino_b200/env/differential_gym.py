"""DifferentialDriveGym: the observation / step side of the reference's gym
wrapper that the planners use (crl_kino/env/differential_gym.py:24-60, 136-145,
180-190).  Reward, reset and rendering are training-time code and out of scope.
"""
import numpy as np

from .. import ops
from .differential_env import DifferentialDriveEnv


class Box:
    """gym.spaces.Box subset: float32 low/high (gym's default dtype)."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is None:
            self.low = np.asarray(low).astype(self.dtype)
            self.high = np.asarray(high).astype(self.dtype)
            self.shape = self.low.shape
        else:
            self.shape = tuple(shape)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)


class DifferentialDriveGym:
    def __init__(self, robot_env=None, reward_param=None, obs_list_list=[]):
        self.robot_env = robot_env if robot_env is not None else \
            DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi)
        self.state_bounds = self.robot_env.get_bounds()["state_bounds"].copy()
        self.state = np.zeros(len(self.state_bounds))
        self.goal = np.zeros(len(self.state))
        self.v_space = Box(low=self.state_bounds[-2:, 0], high=self.state_bounds[-2:, 1])
        self.action_space = self.v_space
        self.sample_positions = self._init_sample_positions()
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(len(self.sample_positions) + 4,))
        self.current_time = 0.0
        self.max_time = 100.0
        self.obs_list_list = obs_list_list

    def _init_sample_positions(self):
        """differential_gym.py:62-75 (the kernel regenerates the same grid)."""
        import itertools
        self.percept_region = np.array([-4, 4, -2, 6])
        self.sample_reso = 0.3
        lr = np.arange(self.percept_region[0], self.percept_region[1], self.sample_reso)
        ba = np.arange(self.percept_region[2], self.percept_region[3], self.sample_reso)
        self.percept_shape = (1, len(ba), len(lr))
        return np.array(list(itertools.product(ba, lr)))

    def _obs(self):
        """differential_gym.py:180-190 -> float64[733]."""
        o64, _, _ = ops.local_obs(self.robot_env.handle, np.asarray(self.state, dtype=np.float64)[None],
                                  np.asarray(self.goal, dtype=np.float64)[None])
        return o64[0].cpu().numpy()

    def obs_batch(self, states, goals, want64=True, want32=False, want_near=False):
        return ops.local_obs(self.robot_env.handle, states, goals, want64, want32, want_near)

    def step(self, action):
        """differential_gym.py:136-145: state update + next observation.  The
        reward / info terms are training-only; info carries what the planners read."""
        dt = 1.0 / 5.0
        self.state = self.robot_env.motion_velocity(self.state, action, dt)
        self.current_time += dt
        obs = self._obs()
        gd = float(np.linalg.norm(self.state[:2] - self.goal[:2]))
        collision = not self.robot_env.valid_state_check(self.state)
        info = {"goal": gd <= 1.0, "goal_dis": gd, "collision": collision}
        done = info["goal"] or collision or self.current_time > self.max_time
        return obs, 0.0, done, info
