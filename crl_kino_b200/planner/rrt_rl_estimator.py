"""RRT_RL_Estimator with the reference's interface
(crl_kino/planner/rrt_rl_estimator.py:17-122): learned parent selection with the
estimator / classifier CNN pair, RL steering inherited from RRT_RL."""
import numpy as np
import torch

from .. import ops
from .rrt_rl import RRT_RL


class RRT_RL_Estimator(RRT_RL):
    def __init__(self, robot_env, policy, estimator, classifier, goal_sample_rate=5, max_iter=500):
        super().__init__(robot_env, policy, goal_sample_rate, max_iter)
        self.estimator = estimator
        self.classifier = classifier

    def choose_parent(self, node_list, rnd_node):
        """rrt_rl_estimator.py:26-79."""
        states = np.array([n.state for n in node_list], dtype=np.float64)
        tgt = np.asarray(rnd_node.state, dtype=np.float64)
        sd = ops.as_dev(states)
        # per-node Euclidean distance on the device (one-node trees share the nearest kernel)
        n = sd.shape[0]
        x = sd[:, 0:1].contiguous()
        y = sd[:, 1:2].contiguous()
        ones = torch.ones(n, dtype=torch.int32, device=sd.device)
        _, dist, _ = ops.nearest(x, y, ones, np.tile(tgt[None, :2], (n, 1)))
        valid = (dist >= 1.0) & (dist < 20.0)
        vidx = torch.nonzero(valid).reshape(-1)
        if vidx.numel() == 0:
            return 0, 100
        goals = ops.as_dev(np.tile(tgt[None], (int(vidx.numel()), 1)))
        _, obs32, _ = ops.local_obs(self.robot_env.handle, sd[vidx].contiguous(), goals,
                                    want64=False, want32=True)
        est = self.estimator.forward_obs(obs32).double()
        prob = self.classifier.forward_obs(obs32).double()
        fail = 1 - torch.round(prob)
        cost = dist.clone()
        cost[vidx] = (est + 1) * 15 + 100 * fail
        c = cost.cpu().numpy()
        i = int(np.argmin(c))
        return i, c[i]
