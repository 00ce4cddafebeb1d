"""DifferentialDriveGym with the reference's interface
(crl_kino/env/differential_gym.py:20-253): observation, step, info, reward and
reset.  Every numeric part runs in libcrlk (crlk_gym_step = motion + _info +
_reward + _obs in one call; clearance tests of reset through crlk_valid_state).
VectorDifferentialDriveGym steps B environments per call (the role
tianshou.env.VectorEnv plays for the reference's training scripts).
Rendering is out of scope.
"""
import itertools

import numpy as np
import torch

from .. import ops
from .differential_env import DifferentialDriveEnv, normalize_angle

INFO_KEYS = ("goal", "goal_dis", "heading", "collision", "clearance", "v", "w", "step")   # dict order of :148-157
REWARD_PARAM = np.array([50.0, -1.5, -0.5, -300.0, 1.0, 1.0, -.5, -2.5])                   # :28-29


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


def _info_dict(row):
    """One row of crlk_gym_step's info matrix -> the reference's info dict (:147-166)."""
    d = {k: float(v) for k, v in zip(INFO_KEYS, row)}
    d["goal"] = bool(row[0] != 0.0)
    d["collision"] = bool(row[3] != 0.0)
    return d


class DifferentialDriveGym:
    def __init__(self, robot_env=None, reward_param=REWARD_PARAM, obs_list_list=[]):
        self.robot_env = robot_env if robot_env is not None else \
            DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi)
        self.state_bounds = self.robot_env.get_bounds()["state_bounds"].copy()
        self.state = np.zeros(len(self.state_bounds))
        self.goal = np.zeros(len(self.state))
        self.v_space = Box(low=self.state_bounds[-2:, 0], high=self.state_bounds[-2:, 1])
        self.action_space = self.v_space
        self.sample_positions = self._init_sample_positions()
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(len(self.sample_positions) + 4,))
        self.n_step = 0
        self.max_step = 200
        self.current_time = 0.0
        self.max_time = 100.0
        self.obs_list_list = obs_list_list
        self.n_case = 0
        self.reward_param = np.array(reward_param, dtype=np.float64).copy()
        self.curriculum = {"obs_num": 7, "ori": False}

    def _init_sample_positions(self):
        """differential_gym.py:62-75 (the kernel regenerates the same grid)."""
        self.percept_region = np.array([-4, 4, -2, 6])
        self.sample_reso = 0.3
        lr = np.arange(self.percept_region[0], self.percept_region[1], self.sample_reso)
        ba = np.arange(self.percept_region[2], self.percept_region[3], self.sample_reso)
        self.percept_shape = (1, len(ba), len(lr))
        return np.array(list(itertools.product(ba, lr)))

    # ---- action <-> velocity space (:110-128) -------------------------------
    @staticmethod
    def _rescale(x, src, dst):
        """Affine map of the box `src` onto the box `dst`, element by element in the
        reference's operation order: (x - lo) / (hi - lo) * (hi' - lo') + lo'."""
        x = np.asarray(x)
        n = len(x)
        out = np.zeros(dst.shape)
        out[:n] = (x - src.low[:n]) / (src.high[:n] - src.low[:n]) * (dst.high[:n] - dst.low[:n]) + dst.low[:n]
        return out

    def a2v(self, action):
        return self._rescale(action, self.action_space, self.v_space)

    def v2a(self, v):
        return self._rescale(v, self.v_space, self.action_space)

    # ---- step / info / reward / obs -----------------------------------------
    def _device_step(self, action):
        s = ops.as_dev(np.asarray(self.state, dtype=np.float64)[None])
        a = None if action is None else np.asarray(action, dtype=np.float64)[None, :2]
        info, reward, o64, _, _ = ops.gym_step(self.robot_env.handle, s, np.asarray(self.goal, dtype=np.float64)[None],
                                               a, self.reward_param, 1.0 / 5.0)
        return s[0].cpu().numpy(), info[0].cpu().numpy(), float(reward[0].item()), o64[0].cpu().numpy()

    def step(self, action):
        """differential_gym.py:136-145."""
        dt = 1.0 / 5.0  # 5 Hz
        self.state, info_row, reward, obs = self._device_step(action)
        self.current_time += dt
        info = _info_dict(info_row)
        done = info["goal"] or info["collision"] or self.current_time > self.max_time
        return obs, reward, done, info

    def _info(self):
        """differential_gym.py:147-166."""
        return _info_dict(self._device_step(None)[1])

    def _reward(self, info):
        """differential_gym.py:168-178."""
        info_arr = np.array([info[a] for a in info])
        return info_arr @ self.reward_param

    def _obs(self):
        """differential_gym.py:180-190 -> float64[733]."""
        o64, _, _ = ops.local_obs(self.robot_env.handle, np.asarray(self.state, dtype=np.float64)[None],
                                  np.asarray(self.goal, dtype=np.float64)[None])
        return o64[0].cpu().numpy()

    def obs_batch(self, states, goals, want64=True, want32=False, want_near=False):
        return ops.local_obs(self.robot_env.handle, states, goals, want64, want32, want_near)

    def set_curriculum(self, **kwargs):
        for k, v in kwargs.items():
            if self.curriculum[k] is not None:
                self.curriculum[k] = v

    def _free(self, state):
        return self.robot_env.get_clearance(state) > 0.5          # the reset rule of :229, :238, :246

    def _draw_goal(self, start, low, high):
        """Up to five polar draws around `start` (:232-239); the last try is kept if all fail."""
        (x0, x1), (y0, y1) = self.state_bounds[0], self.state_bounds[1]
        goal = np.zeros(len(self.state_bounds))
        for _ in range(5):
            r = np.random.uniform(low, high)
            theta = np.random.uniform(-np.pi, np.pi)
            goal[0] = np.clip(start[0] + r * np.cos(theta), x0, x1)
            goal[1] = np.clip(start[1] + r * np.sin(theta), y0, y1)
            if self._free(goal):
                break
        return goal

    def reset(self, low=5, high=8):
        """differential_gym.py:215-253.  Consumes numpy's global RNG with the same calls in the
        same order as the reference, so a seeded episode starts exactly like the reference's."""
        assert len(self.obs_list_list) > 0, "No defined training environments"
        self.robot_env.set_obs(self.obs_list_list[np.random.randint(0, len(self.obs_list_list))])
        lo, hi = self.state_bounds[:, 0], self.state_bounds[:, 1]
        while True:
            start = np.random.uniform(lo, hi)
            if not self._free(start):
                continue
            goal = self._draw_goal(start, low, high)
            if self.curriculum["ori"]:          # face the goal (:242-244)
                start[2] = normalize_angle(np.arctan2(goal[1] - start[1], goal[0] - start[0]))
            gap = np.linalg.norm(start[:2] - goal[:2])
            if self._free(start) and self._free(goal) and low < gap < high:
                break
        self.state, self.goal = start, goal
        self.current_time = 0
        return self._obs()


class VectorDifferentialDriveGym:
    """B DifferentialDriveGym environments over ONE obstacle set, stepped by one
    crlk_gym_step call (states, goals and clocks stay on the device).  step() returns
    CUDA tensors: obs [B, 733] float64 (or float32 rows with obs32=True), reward [B],
    done [B] bool, info [B, 8] (columns = INFO_KEYS).  Finished environments are NOT
    reset implicitly; call reset(mask) like tianshou's VectorEnv.reset(id)."""

    def __init__(self, robot_env, n_envs, reward_param=REWARD_PARAM, seed=0, obs32=False):
        self.robot_env = robot_env
        self.n = int(n_envs)
        self.reward_param = np.array(reward_param, dtype=np.float64).copy()
        self.state_bounds = robot_env.get_bounds()["state_bounds"].copy()
        d = ops.dev()
        self.states = torch.zeros(self.n, 5, dtype=torch.float64, device=d)
        self.goals = torch.zeros(self.n, 5, dtype=torch.float64, device=d)
        self.current_time = torch.zeros(self.n, dtype=torch.float64, device=d)
        self.max_time = 100.0
        self.rng = np.random.default_rng(seed)
        self.obs32 = obs32
        self.curriculum = {"obs_num": 7, "ori": False}

    def _observe(self, actions):
        info, reward, o64, o32, _ = ops.gym_step(self.robot_env.handle, self.states, self.goals, actions,
                                                 self.reward_param, 1.0 / 5.0, want64=not self.obs32,
                                                 want32=self.obs32)
        return info, reward, (o32 if self.obs32 else o64)

    def step(self, actions):
        """actions: [B, 2] velocity commands (v, w) -> (obs, reward, done, info)."""
        info, reward, obs = self._observe(ops.as_dev(actions).reshape(self.n, 2))
        self.current_time += 1.0 / 5.0
        done = (info[:, 0] != 0) | (info[:, 3] != 0) | (self.current_time > self.max_time)
        return obs, reward, done, info

    def set_states(self, states, goals):
        self.states.copy_(ops.as_dev(states).reshape(self.n, 5))
        self.goals.copy_(ops.as_dev(goals).reshape(self.n, 5))
        self.current_time.zero_()
        return self._observe(None)[2]

    def reset(self, mask=None, low=5, high=8):
        """The reference's reset rule (:215-253) for the environments selected by `mask`
        (default: all), drawn in batches: start uniform in the state bounds with clearance > 0.5,
        goal at distance r ~ U(low, high) in a uniform direction (clipped to the bounds, up to 5
        tries) with clearance > 0.5 and low < ||start - goal|| < high.  Uses this object's own
        Generator (the reference's global-RNG order is defined for one environment only)."""
        sel = np.arange(self.n) if mask is None else np.flatnonzero(np.asarray(
            mask.cpu().numpy() if isinstance(mask, torch.Tensor) else mask))
        todo = sel.copy()
        sb = self.state_bounds
        new_s = np.zeros((len(sel), 5))
        new_g = np.zeros((len(sel), 5))
        slot = {int(e): i for i, e in enumerate(sel)}
        env = self.robot_env

        def clear(x):
            return env.valid_state_batch(x, want_clearance=True)[1].cpu().numpy()
        guard = 0
        while len(todo) and guard < 1000:
            guard += 1
            m = len(todo)
            start = self.rng.uniform(sb[:, 0], sb[:, 1], (m, 5))
            ok_s = clear(start) > 0.5
            goal = np.zeros((m, 5))
            got = np.zeros(m, dtype=bool)
            for _ in range(5):
                r = self.rng.uniform(low, high, m)
                th = self.rng.uniform(-np.pi, np.pi, m)
                cand = np.zeros((m, 5))
                cand[:, 0] = np.clip(start[:, 0] + r * np.cos(th), sb[0, 0], sb[0, 1])
                cand[:, 1] = np.clip(start[:, 1] + r * np.sin(th), sb[1, 0], sb[1, 1])
                goal[~got] = cand[~got]                 # the reference keeps the last try when all five fail
                got |= ~got & (clear(cand) > 0.5)
            if self.curriculum["ori"]:
                ang = np.arctan2(goal[:, 1] - start[:, 1], goal[:, 0] - start[:, 0])
                start[:, 2] = [normalize_angle(a) for a in ang]
                ok_s = clear(start) > 0.5
            dist = np.linalg.norm(start[:, :2] - goal[:, :2], axis=1)
            ok = ok_s & got & (clear(goal) > 0.5) & (low < dist) & (dist < high)
            for e, s_, g_ in zip(todo[ok], start[ok], goal[ok]):
                new_s[slot[int(e)]] = s_
                new_g[slot[int(e)]] = g_
            todo = todo[~ok]
        if len(todo):
            raise RuntimeError("reset: no valid start/goal configuration found")
        idx = torch.as_tensor(sel, device=self.states.device)
        self.states[idx] = ops.as_dev(new_s)
        self.goals[idx] = ops.as_dev(new_g)
        self.current_time[idx] = 0.0
        return self._observe(None)[2]
