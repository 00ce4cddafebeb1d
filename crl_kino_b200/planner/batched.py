"""Batched, device-resident form of the planner's hot path: Q independent
queries, each with its own tree (SoA node tables in HBM), advanced together.

BatchedExtender.step() is one pass of the reference's loop body for every query
(rrt.py:68-72): nearest node of the query's sample (crlk_nearest), pick that
node's state (crlk_gather_rows), one whole extension toward the sample
(crlk_extend_dwa or crlk_extend_rl).  PyTorch only owns the buffers.
"""
import numpy as np
import torch

from .. import ops


class Forest:
    """Q trees in SoA layout: x/y float64 [Q, cap] (+ float32 copies that the
    nearest-node scan streams), full states float64 [Q, cap, 5], parent int32
    [Q, cap], n_nodes int32 [Q]."""

    def __init__(self, Q, capacity, device=None, with_f32=True):
        d = device or ops.dev()
        self.Q, self.cap = Q, capacity
        self.x = torch.zeros(Q, capacity, dtype=torch.float64, device=d)
        self.y = torch.zeros(Q, capacity, dtype=torch.float64, device=d)
        self.x32 = torch.zeros(Q, capacity, dtype=torch.float32, device=d) if with_f32 else None
        self.y32 = torch.zeros(Q, capacity, dtype=torch.float32, device=d) if with_f32 else None
        self.states = torch.zeros(Q, capacity, 5, dtype=torch.float64, device=d)
        self.parent = torch.full((Q, capacity), -1, dtype=torch.int32, device=d)
        self.n_nodes = torch.zeros(Q, dtype=torch.int32, device=d)

    def fill(self, states, n_nodes=None):
        """Bulk-load node states [Q, N, 5] (setup path, not timed)."""
        s = ops.as_dev(states)
        N = s.shape[1]
        self.states[:, :N] = s
        self.x[:, :N] = s[:, :, 0]
        self.y[:, :N] = s[:, :, 1]
        if self.x32 is not None:
            self.x32[:, :N] = s[:, :, 0].float()
            self.y32[:, :N] = s[:, :, 1].float()
        self.n_nodes[:] = N if n_nodes is None else ops.as_dev(n_nodes, torch.int32)

    def bytes_scanned_per_query(self):
        return 8 if self.x32 is not None else 16


class BatchedExtender:
    """nearest -> gather -> extend for Q queries per call."""

    def __init__(self, env, forest, dwa=None, policy=None, eps=0.0, t_max_extend=10.0, t_tree=5.0):
        self.env = env
        self.forest = forest
        self.dwa = dwa
        self.policy = policy
        self.eps = eps
        self.xp = ops.extend_params(t_max_extend, t_tree)
        self.coord_bound = env.coord_bound()
        Q = forest.Q
        d = forest.x.device
        self.from_states = torch.empty(Q, 5, dtype=torch.float64, device=d)
        self.stats = torch.zeros(4, dtype=torch.int64, device=d)
        self.out = None
        self.max_nodes = forest.cap
        # pinned staging for the host-buffer entry point
        self._h_samples = torch.empty(Q, 5, dtype=torch.float64).pin_memory()
        self._d_samples = torch.empty(Q, 5, dtype=torch.float64, device=d)
        per = self.xp.n_max_extend // self.xp.n_tree + 1
        self._h_out = dict(path=torch.empty(Q, self.xp.n_max_extend, 5, dtype=torch.float64).pin_memory(),
                           n_steps=torch.empty(Q, dtype=torch.int32).pin_memory(),
                           status=torch.empty(Q, dtype=torch.int32).pin_memory(),
                           node_step=torch.empty(Q, per, dtype=torch.int32).pin_memory(),
                           n_nodes=torch.empty(Q, dtype=torch.int32).pin_memory(),
                           nearest_idx=torch.empty(Q, dtype=torch.int32).pin_memory(),
                           flags=torch.empty(Q, dtype=torch.uint8).pin_memory())
        self.launches_per_step = 4 if policy is None else 3 + self.xp.n_max_extend * (2 + len(policy.actor.weights)) + 1

    def step(self, samples, events=None):
        """samples: float64 CUDA [Q, 5].  Returns the extension outputs (CUDA tensors)."""
        f = self.forest
        ev = events

        def mark(i):
            if ev is not None:
                ev[i].record()
        mark(0)
        idx, dist, tie = ops.nearest(f.x, f.y, f.n_nodes, samples, x32=f.x32, y32=f.y32,
                                     coord_bound=self.coord_bound, max_nodes=self.max_nodes)
        mark(1)
        ops.gather_rows(f.states, idx, out=self.from_states)
        mark(2)
        if self.policy is None:
            self.out = ops.extend_dwa(self.env.handle, self.dwa.params(), self.xp, self.from_states, samples,
                                      out=self.out, stats=self.stats)
        else:
            self.out = ops.extend_rl(self.env.handle, self.policy.actor.handle, self.xp, self.from_states,
                                     samples, None, self.eps)
        mark(3)
        self.out["nearest_idx"], self.out["nearest_dist"] = idx, dist
        return self.out

    def step_host(self, samples_np):
        """Host-buffer entry point: numpy [Q, 5] in, dict of numpy views out
        (pinned staging; H2D and D2H copies are part of the call)."""
        self._h_samples.numpy()[...] = samples_np
        self._d_samples.copy_(self._h_samples, non_blocking=True)
        out = self.step(self._d_samples)
        for k, h in self._h_out.items():
            h.copy_(out[k], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return {k: h.numpy() for k, h in self._h_out.items()}

    def h2d_bytes(self):
        return self._h_samples.numel() * 8

    def d2h_bytes(self):
        return sum(h.numel() * h.element_size() for h in self._h_out.values())
