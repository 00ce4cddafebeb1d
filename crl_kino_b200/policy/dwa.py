"""DWA controller with the reference's interface (crl_kino/policy/dwa.py:12-86);
the rollout grid, scoring and argmin run in crlk_dwa_control."""
import numpy as np

from .. import ops


class DWA:
    def __init__(self, robot_env, v_res=0.01, w_res=0.1, dt=0.1, predict_time=1.0,
                 to_goal_cost_gain=1.0, ob_gain=1.0, speed_cost_gain=0.3, skip=3, use_clearance=False):
        self.robot_env = robot_env
        self.v_res = v_res
        self.w_res = w_res
        self.dt = dt
        self.predict_time = predict_time
        self.to_goal_cost_gain = to_goal_cost_gain
        self.speed_cost_gain = speed_cost_gain
        self.ob_gain = ob_gain
        self.skip = skip
        # False = the shipped reference (dwa.py:64-69 commented out);
        # True = those lines evaluated (per-rollout clearance scoring).
        self.use_clearance = use_clearance

    def set_dwa(self, **kwargs):
        for key, value in kwargs.items():
            if key in ("v_res", "w_res", "dt", "predict_time", "to_goal_cost_gain", "ob_gain",
                       "speed_cost_gain", "skip", "use_clearance"):
                setattr(self, key, value)

    def params(self):
        return ops.dwa_params(self.v_res, self.w_res, self.dt, self.predict_time,
                              self.to_goal_cost_gain, self.ob_gain, self.speed_cost_gain, self.skip,
                              self.use_clearance)

    def calc_trajectory(self, x_init, v):
        """dwa.py:35-43."""
        h = self.robot_env.handle
        n_rows = ops.dwa_traj_rows(h, self.params())
        x = np.asarray(x_init, dtype=np.float64)
        traj = [x.copy()]
        cmd = np.asarray(v, dtype=np.float64)[None, :2]
        s = ops.as_dev(x[None])
        for _ in range(n_rows - 1):
            s = ops.motion_velocity(h, s, cmd, self.dt)
            traj.append(s[0].cpu().numpy())
        return np.vstack(traj)

    def control(self, x_init, goal):
        """dwa.py:45-86 -> (opt_v[2], opt_traj[rows, 5])."""
        r = ops.dwa_control(self.robot_env.handle, self.params(),
                            np.asarray(x_init, dtype=np.float64)[None],
                            np.asarray(goal, dtype=np.float64)[None], want_traj=True)
        return r["opt_v"][0].cpu().numpy(), r["opt_traj"][0].cpu().numpy()

    def control_batch(self, states, goals, want_traj=False):
        """B independent decisions; returns the dict of CUDA tensors of ops.dwa_control."""
        return ops.dwa_control(self.robot_env.handle, self.params(), states, goals, want_traj)
