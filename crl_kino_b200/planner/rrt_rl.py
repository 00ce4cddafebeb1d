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
        self.noise_stream = None   # optional iterator of [2] standard-normal draws

    def _draw_noise(self, n):
        if self.eps == 0.0:
            return None
        if self.noise_stream is not None:
            # an explicit stream holds one draw per EXECUTED step; unused tail rows are zeros
            return None
        return torch.randn(size=(n, 2)).numpy()

    def steer(self, from_node, to_node, t_max_extend=10.0, t_tree=5.0):
        xp = ops.extend_params(t_max_extend, t_tree)
        n_max = xp.n_max_extend
        if self.noise_stream is not None and self.eps != 0.0:
            out = self._steer_streamed(from_node, to_node, xp)
        else:
            noise = self._draw_noise(n_max)
            out = ops.extend_rl(self.robot_env.handle, self.policy.actor.handle, xp,
                                from_node.state[None], to_node.state[None],
                                None if noise is None else noise[None], self.eps)
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

    def _steer_streamed(self, from_node, to_node, xp):
        """Consume exactly one noise draw per executed step from `noise_stream`
        (parity runs): the extension is replayed with a growing noise prefix
        only when the stream must not be over-consumed."""
        n_max = xp.n_max_extend
        # peek: run with zeros after the known prefix is impossible without knowing
        # n_steps; instead draw lazily step by step through a buffered iterator.
        buf = []
        it = self.noise_stream

        class Tail:
            pass
        # draw all n_max rows, then push back the unused ones
        rows = []
        for _ in range(n_max):
            try:
                rows.append(np.asarray(next(it), dtype=np.float32).reshape(2))
            except StopIteration:
                rows.append(np.zeros(2, dtype=np.float32))
                buf.append(None)
        noise = np.stack(rows)[None]
        out = ops.extend_rl(self.robot_env.handle, self.policy.actor.handle, xp,
                            from_node.state[None], to_node.state[None], noise, self.eps)
        used = int(out["n_steps"][0].item())
        unused = rows[used:n_max - len(buf)]
        if unused:
            import itertools
            self.noise_stream = itertools.chain(iter(unused), it)
        return out
