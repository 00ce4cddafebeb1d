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
        """rrt_rl_estimator.py:26-79: distance filter, observations, both CNNs, cost and
        argmin in one device call (crlk_estimator_choose_parent)."""
        states = ops.as_dev(np.array([n.state for n in node_list], dtype=np.float64))[None].contiguous()
        n = torch.tensor([len(node_list)], dtype=torch.int32, device=states.device)
        idx, cost = ops.estimator_choose_parent(self.robot_env.handle, self.estimator.pair.handle, states, n,
                                                np.asarray(rnd_node.state, dtype=np.float64)[None])
        return int(idx[0].item()), float(cost[0].item())
