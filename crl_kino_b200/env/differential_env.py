"""DifferentialDriveEnv with the reference's interface
(crl_kino/env/differential_env.py:27-169); all numerics run on the GPU through
libcrlk (crlk_motion_velocity, crlk_valid_state, crlk_valid_points, ...).

Single-state methods keep the reference signatures and return numpy values;
the *_batch methods take [B, ...] arrays or CUDA tensors and return CUDA
tensors.
"""
import math

import numpy as np
import torch

from .. import ops
from ..utils.rigid import CircleRobot, Obstacle, RectRobot, rects_of
from .robot_env import RobotEnv


def normalize_angle(angle):
    """differential_env.py:20-24 (host helper; the kernels carry their own)."""
    norm_angle = angle % (2 * math.pi)
    if norm_angle > math.pi:
        norm_angle -= 2 * math.pi
    return norm_angle


class DifferentialDriveEnv(RobotEnv):
    def __init__(self, max_v=1.0, min_v=-0.1, max_w=np.pi, max_acc_v=1.0, max_acc_w=np.pi,
                 env_bounds=[-20, 20, -20, 20], base_dt=0.1):
        super().__init__()
        self.max_v = max_v
        self.min_v = min_v
        self.max_w = max_w
        self.min_w = -self.max_w
        self.max_acc_v = max_acc_v
        self.min_acc_v = -self.max_acc_v
        self.max_acc_w = max_acc_w
        self.min_acc_w = -self.max_acc_w
        self._handle = None
        self.rigid_robot = RectRobot(np.array([[-0.7, -0.5], [0.7, 0.5]]), color="k")
        self.env_bounds = np.zeros((2, 2))
        self.env_bounds[0, 0] = env_bounds[0]
        self.env_bounds[0, 1] = env_bounds[1]
        self.env_bounds[1, 0] = env_bounds[2]
        self.env_bounds[1, 1] = env_bounds[3]
        self._obs_list = []
        self.base_dt = base_dt
        self._handle = None
        self._rects = np.zeros((0, 4), dtype=np.float32)
        self._dirty = True

    # ---- footprint (differential_env.py:41; RectRobot or CircleRobot, rigid.py:256-347) ----
    @property
    def rigid_robot(self):
        return self._rigid_robot

    @rigid_robot.setter
    def rigid_robot(self, robot):
        if not isinstance(robot, (RectRobot, CircleRobot)):
            raise NotImplementedError("Unsupported robot footprint")
        self._rigid_robot = robot
        self._handle = None          # the footprint is baked into the device handle

    # ---- device handle -------------------------------------------------
    @property
    def handle(self):
        """The CrlkEnv handle (created lazily on the current CUDA device)."""
        if self._handle is None:
            rb = self._rigid_robot
            circle = isinstance(rb, CircleRobot)
            rect = np.array([-0.7, -0.5, 0.7, 0.5], dtype=np.float32) if circle else rb.rect.reshape(4)
            self._handle = ops.EnvHandle(self.max_v, self.min_v, self.max_w, self.max_acc_v,
                                         self.max_acc_w, self.base_dt, self.env_bounds.reshape(4),
                                         self._rects, rect, float(rb.radius) if circle else 0.0)
            self._dirty = False
        elif self._dirty:
            self._handle.set_obs(self._rects)
            self._dirty = False
        return self._handle

    @property
    def rects(self):
        return self._rects

    def coord_bound(self):
        """Upper bound of |coordinate| of anything inside the map (walls included)."""
        b = float(np.max(np.abs(self.env_bounds)))
        if len(self._rects):
            b = max(b, float(np.max(np.abs(self._rects))))
        return b

    # ---- obstacles (differential_env.py:116-123) -------------------------
    # The flat float32 [M, 4] table is the source of truth (it is what the device holds); the
    # reference's list of RectObs objects is built from it on demand.
    @property
    def obs_list(self):
        if self._obs_list is None:
            from ..utils.rigid import RectObs
            self._obs_list = [RectObs(r.reshape(2, 2)) for r in self._rects]
        return self._obs_list

    @obs_list.setter
    def obs_list(self, obs_list):
        self.set_obs(obs_list)

    def add_obs(self, obs):
        assert isinstance(obs, Obstacle), "Unsupported obstacle type"
        if self._obs_list is not None:
            self._obs_list.append(obs)
        self._rects = np.ascontiguousarray(np.vstack([self._rects, rects_of([obs])]), dtype=np.float32)
        self._dirty = True

    def set_obs(self, obs_list):
        if isinstance(obs_list, np.ndarray):       # flat float32 [M, 4] fast path
            self._rects = rects_of(obs_list)
            self._obs_list = None                  # built lazily (a dense map has 10^4 rectangles)
            self._dirty = True
            return
        for obs in obs_list:
            assert isinstance(obs, Obstacle), "Unsupported obstacle type"
        self._obs_list = list(obs_list)
        self._rects = rects_of(self._obs_list)
        self._dirty = True

    # ---- dynamics ----------------------------------------------------------
    def motion_velocity(self, state, velocity, duration):
        """differential_env.py:55-73."""
        out = ops.motion_velocity(self.handle, np.asarray(state, dtype=np.float64)[None],
                                  np.asarray(velocity, dtype=np.float64)[None, :2], duration)
        return out[0].cpu().numpy()

    def motion(self, state, input_u, duration):
        """differential_env.py:76-83."""
        out = ops.motion_velocity(self.handle, np.asarray(state, dtype=np.float64)[None],
                                  np.asarray(input_u, dtype=np.float64)[None, :2], duration, accel=True)
        return out[0].cpu().numpy()

    def motion_base(self, state, input_u, dt):
        """differential_env.py:85-105: one sub-step; mutates and returns `state`."""
        if dt != self.base_dt:
            # the device integrates in base_dt sub-steps (the only way the reference itself calls
            # motion_base, differential_env.py:71,81); other step sizes are not offered
            raise NotImplementedError("motion_base is only available at dt == base_dt")
        out = ops.motion_velocity(self.handle, np.asarray(state, dtype=np.float64)[None],
                                  np.asarray(input_u, dtype=np.float64)[None, :2], dt, accel=True)
        state[:] = out[0].cpu().numpy()
        return state

    def dynamic_window(self, state, dt):
        """differential_env.py:107-114 -> [[v_min, v_max], [w_min, w_max]]."""
        out = ops.dynamic_window(self.handle, np.asarray(state, dtype=np.float64)[None], dt)
        return out[0].cpu().numpy().reshape(2, 2)

    def motion_velocity_batch(self, states, velocities, duration):
        return ops.motion_velocity(self.handle, states, velocities, duration)

    # ---- validity ------------------------------------------------------------
    def get_clearance(self, state):
        """differential_env.py:126-132."""
        self.rigid_robot.pose = np.asarray(state, dtype=float)[:3]
        _, clr, _ = ops.valid_state(self.handle, np.asarray(state, dtype=np.float64)[None],
                                    want_clearance=True)
        return float(clr[0].item())

    def valid_point_check(self, points):
        """differential_env.py:135-144."""
        v = ops.valid_points(self.handle, np.asarray(points, dtype=np.float64).reshape(-1, 2))
        return v.cpu().numpy().astype(bool)

    def valid_state_check(self, state):
        """differential_env.py:146-151."""
        self.rigid_robot.pose = np.asarray(state, dtype=float)[:3]
        v, _, _ = ops.valid_state(self.handle, np.asarray(state, dtype=np.float64)[None])
        return bool(v[0].item())

    def valid_state_batch(self, states, want_clearance=False, want_near=False):
        """-> (valid u8[B], clearance f64[B] | None, near_boundary u8[B] | None) CUDA tensors."""
        return ops.valid_state(self.handle, states, want_clearance, want_near)

    def get_bounds(self):
        """differential_env.py:153-169."""
        state_bounds = np.zeros((5, 2))
        state_bounds[:2, :] = self.env_bounds
        state_bounds[2, 0] = -np.pi
        state_bounds[2, 1] = np.pi
        state_bounds[3, 0] = self.min_v
        state_bounds[3, 1] = self.max_v
        state_bounds[4, 0] = self.min_w
        state_bounds[4, 1] = self.max_w
        control_bounds = np.zeros((2, 2))
        control_bounds[0, 0] = self.min_acc_v
        control_bounds[0, 1] = self.max_acc_v
        control_bounds[1, 0] = self.min_acc_w
        control_bounds[1, 1] = self.max_acc_w
        return {"state_bounds": state_bounds, "control_bounds": control_bounds}
