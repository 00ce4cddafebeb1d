"""RRT with the reference's interface (crl_kino/planner/rrt.py:21-161).

`planning()` keeps the reference's host loop (sampling with Python's `random` /
`numpy.random`, tree bookkeeping); `steer` runs one whole extension -- up to 50
DWA control steps with motion and collision checks -- in a single kernel launch
(crlk_extend_dwa) and `choose_parent` scans the tree on the device
(crlk_nearest).  BatchedRRT (planner/batched.py) advances many queries at once.
"""
import random
import time
from abc import ABC

import numpy as np
import torch

from .. import ops
from ..policy.dwa import DWA


class RRT(ABC):
    class Node:
        def __init__(self, state):
            self.state = np.array(state, dtype=np.float64).copy()
            self.path = []
            self.parent = None

    def __init__(self, robot_env, goal_sample_rate=5, max_iter=500):
        self.robot_env = robot_env
        self.dwa = DWA(robot_env)
        self.dwa.set_dwa(dt=0.2, v_res=0.02)
        self.goal_sample_rate = goal_sample_rate
        self.max_iter = max_iter
        self.node_list = []
        self.start = None
        self.goal = None
        self.planning_time = 0.0
        self.path = None
        self.state_bounds = self.robot_env.get_bounds()["state_bounds"]
        self.verbose = True
        self.n_extend_calls = 0
        self.n_control_steps = 0
        self.flags_seen = 0          # OR of the extensions' report bits (near-boundary, near-tie)

    def planning(self):
        """rrt.py:51-89."""
        if self.start is None or self.goal is None:
            raise ValueError("start or goal is not set")
        tic = time.time()
        self.node_list = [self.start]
        reach_exactly = False
        for i in range(self.max_iter):
            if self.verbose:
                print("Iteration " + str(i) + ":", str(len(self.node_list)) + " nodes")
            good_sample = False
            while not good_sample:
                rnd_node = self.sample(self.goal)
                if not self.robot_env.valid_state_check(rnd_node.state):
                    continue
                nearest_ind, cost = self.choose_parent(self.node_list, rnd_node)
                if 1.0 <= cost < 20:
                    good_sample = True
            nearest_node = self.node_list[nearest_ind]
            new_node_list = self.steer(nearest_node, rnd_node)
            if len(new_node_list) > 0:
                if self._dist(self.node_list[-1].state, self.goal.state) <= 1.0:
                    reach_exactly = True
                    break
                nearest_ind, cost = self.choose_parent(new_node_list, self.goal)
                if cost < 5.0:
                    new_node_list = self.steer(new_node_list[nearest_ind], self.goal)
                    if self._dist(self.node_list[-1].state, self.goal.state) <= 1.0:
                        reach_exactly = True
                        break
        path = self.generate_final_course(-1)
        self.planning_time = time.time() - tic
        self.path = path
        self.reach_exactly = reach_exactly
        return path

    @staticmethod
    def _dist(a, b):
        """np.linalg.norm over (x, y) evaluated on the device (one-node nearest)."""
        _, d = RRT._nearest_xy(np.array([[a[0], a[1]]]), b)
        return d

    def set_start_and_goal(self, start, goal):
        assert self.robot_env.valid_state_check(start) and self.robot_env.valid_state_check(goal), \
            "The start or goal states are not valid"
        self.start = self.Node(start)
        self.goal = self.Node(goal)
        self._check_window(self.start.state)

    def _check_window(self, state):
        """A state whose v or omega lies outside the vehicle limits by more than acc * dt has an
        empty DWA grid; the reference fails on it inside DWA.control (dwa.py:76).  Fail early."""
        e, dt = self.robot_env, self.dwa.dt
        v, w = float(state[3]), float(state[4])
        if (min(v + e.max_acc_v * dt, e.max_v) + self.dwa.v_res <= max(v - e.max_acc_v * dt, e.min_v) or
                min(w + e.max_acc_w * dt, e.max_w) + self.dwa.w_res <= max(w - e.max_acc_w * dt, -e.max_w)):
            raise ValueError("state velocity outside the limits: the dynamic window is empty")

    def _segments_to_nodes(self, from_node, path, n_steps, node_steps, first_path_has_start):
        """Rebuild the reference's Node objects from an extension's flat output."""
        new_node_list = []
        parent = from_node
        prev = 0
        for k, stp in enumerate(node_steps):
            node = self.Node(path[stp - 1])
            node.parent = parent
            seg = [path[i].copy() for i in range(prev, stp)]
            if k == 0 and not first_path_has_start:
                node.path = seg                                  # rrt.py:102-104
            else:
                node.path = [parent.state.copy()] + seg          # rrt.py:133 / rrt_rl.py:27
            node.state = seg[-1]
            new_node_list.append(node)
            parent = node
            prev = stp
        return new_node_list

    def steer(self, from_node, to_node, t_max_extend=10.0, t_tree=5.0):
        """rrt.py:99-135, DWA steering; one kernel launch per extension."""
        xp = ops.extend_params(t_max_extend, t_tree)
        out = ops.extend_dwa(self.robot_env.handle, self.dwa.params(), xp,
                             from_node.state[None], to_node.state[None])
        n_steps = int(out["n_steps"][0].item())
        nn = int(out["n_nodes"][0].item())
        path = out["path"][0, :n_steps].cpu().numpy()
        node_steps = out["node_step"][0, :nn].cpu().numpy().tolist()
        self.n_extend_calls += 1
        self.n_control_steps += n_steps
        self.last_extend = dict(n_steps=n_steps, status=int(out["status"][0].item()),
                                flags=int(out["flags"][0].item()))
        self.flags_seen |= self.last_extend["flags"]
        if self.last_extend["status"] == 3:          # CRLK_EXT_INVALID
            raise ValueError("state velocity outside the limits: the dynamic window is empty")
        new_node_list = self._segments_to_nodes(from_node, path, n_steps, node_steps, False)
        self.node_list += new_node_list
        return new_node_list

    def generate_final_course(self, goal_ind):
        """rrt.py:137-144."""
        path = []
        node = self.node_list[goal_ind]
        while node.parent is not None:
            path = node.path[1:] + path
            node = node.parent
        return [node.state] + path

    def sample(self, goal):
        """rrt.py:146-154 (host RNG, same call order as the reference)."""
        if random.randint(0, 100) > self.goal_sample_rate:
            state = np.random.uniform(self.state_bounds[:, 0], self.state_bounds[:, 1])
            rnd = self.Node(state)
        else:
            rnd = self.Node(self.goal.state)
        return rnd

    @staticmethod
    def _nearest_xy(xy, target):
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        x = ops.as_dev(xy[:, 0].copy())[None]
        y = ops.as_dev(xy[:, 1].copy())[None]
        n = torch.tensor([xy.shape[0]], dtype=torch.int32, device=x.device)
        idx, dist, _ = ops.nearest(x, y, n, np.asarray(target, dtype=np.float64)[None, :2])
        return int(idx[0].item()), float(dist[0].item())

    @staticmethod
    def choose_parent(node_list, rnd_node):
        """rrt.py:156-161: nearest node in (x, y), first minimum."""
        xy = np.array([n.state[:2] for n in node_list])
        return RRT._nearest_xy(xy, rnd_node.state)

    def collsion_check(self, path):
        v, _, _ = self.robot_env.valid_state_batch(np.asarray(path, dtype=np.float64))
        return bool(v.all().item())
