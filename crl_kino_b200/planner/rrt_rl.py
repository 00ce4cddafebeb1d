"""RRT_RL with the reference's interface (crl_kino/planner/rrt_rl.py:15-63):
the DDPG policy as steering function.  One `steer` = one crlk_extend_rl call
(observation kernel -> actor MLP -> motion + collision, up to 50 steps)."""
import numpy as np
import torch

from .. import ops
from ..env import DifferentialDriveGym
from .rrt import RRT


class RRT_RL(RRT):
    def __init__(self, robot_env, policy, goal_sample_rate=5, max_iter=500):
        super().__init__(robot_env, goal_sample_rate, max_iter)
        self.gym = DifferentialDriveGym(robot_env)
        self.policy = policy
        self.eps = 0.05            # rrt_rl.py:41
        self._noise = None         # optional explicit stream of standard-normal draws [S, 2]
        self._cursor = None

    def set_noise_stream(self, draws):
        """Use these standard-normal draws (one row per executed policy step, in order)
        instead of torch.randn -- the explicit form of the noise the reference draws
        inside the actor (net.py:147-150)."""
        self._noise = ops.as_dev(np.asarray(draws, dtype=np.float32).reshape(1, -1, 2), torch.float32)
        self._cursor = torch.zeros(1, dtype=torch.int32, device=self._noise.device)

    def steer(self, from_node, to_node, t_max_extend=10.0, t_tree=5.0):
        """rrt_rl.py:22-63: one crlk_extend_rl call."""
        xp = ops.extend_params(t_max_extend, t_tree)
        noise, cursor = None, None
        if self.eps != 0.0:
            if self._noise is not None:
                noise, cursor = self._noise, self._cursor
                if self.n_control_steps + xp.n_max_extend > self._noise.shape[1]:
                    raise RuntimeError("the exploration noise stream is too short for another extension: "
                                       f"{self.n_control_steps} draws used of {self._noise.shape[1]}")
            else:
                noise = torch.randn(size=(1, xp.n_max_extend, 2)).numpy()
        out = ops.extend_rl(self.robot_env.handle, self.policy.actor.handle, xp, from_node.state[None],
                            to_node.state[None], noise, self.eps, noise_cursor=cursor)
        n_steps = int(out["n_steps"][0].item())
        nn = int(out["n_nodes"][0].item())
        path = out["path"][0, :n_steps].cpu().numpy()
        node_steps = out["node_step"][0, :nn].cpu().numpy().tolist()
        self.n_extend_calls += 1
        self.n_control_steps += n_steps
        self.last_extend = dict(n_steps=n_steps, status=int(out["status"][0].item()),
                                flags=int(out["flags"][0].item()))
        new_node_list = self._segments_to_nodes(from_node, path, n_steps, node_steps, True)
        self.node_list += new_node_list
        return new_node_list
