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
    """Q trees in HBM: full states float64 [Q, cap, 5] (the exact coordinates are its columns 0 / 1),
    float32 copies x32 / y32 [Q, cap] that the nearest-node scan streams (8 B per node), parent
    int32 [Q, cap], n_nodes int32 [Q].  with_xy64 adds separate float64 x / y arrays [Q, cap] (the
    SoA view the device-resident planners index, crlk_rrt_append / crlk_plan_dwa); without them a
    node costs 48 + 4 bytes instead of 64 + 4, which is what lets cfg5's 100,000-node trees of
    32,768 queries fit one 180 GB GPU."""

    def __init__(self, Q, capacity, device=None, with_f32=True, with_xy64=True, with_parent=True):
        d = device or ops.dev()
        self.Q, self.cap = Q, capacity
        self.x = torch.zeros(Q, capacity, dtype=torch.float64, device=d) if with_xy64 else None
        self.y = torch.zeros(Q, capacity, dtype=torch.float64, device=d) if with_xy64 else None
        self.x32 = torch.zeros(Q, capacity, dtype=torch.float32, device=d) if with_f32 else None
        self.y32 = torch.zeros(Q, capacity, dtype=torch.float32, device=d) if with_f32 else None
        self.states = torch.zeros(Q, capacity, 5, dtype=torch.float64, device=d)
        self.parent = torch.full((Q, capacity), -1, dtype=torch.int32, device=d) if with_parent else None
        self.n_nodes = torch.zeros(Q, dtype=torch.int32, device=d)

    def fill(self, states, n_nodes=None):
        """Bulk-load node states [Q, N, 5] (setup path, not timed)."""
        s = ops.as_dev(states)
        N = s.shape[1]
        self.states[:, :N] = s
        if self.x is not None:
            self.x[:, :N] = s[:, :, 0]
            self.y[:, :N] = s[:, :, 1]
        if self.x32 is not None:
            self.x32[:, :N] = s[:, :, 0].float()
            self.y32[:, :N] = s[:, :, 1].float()
        self.n_nodes[:] = N if n_nodes is None else ops.as_dev(n_nodes, torch.int32)

    def nearest(self, targets, coord_bound, max_nodes=None):
        """RRT.choose_parent (rrt.py:156-161) for one target per tree -> (idx, dist, near_tie)."""
        if self.x is not None:
            return ops.nearest(self.x, self.y, self.n_nodes, targets, x32=self.x32, y32=self.y32,
                               coord_bound=coord_bound, max_nodes=max_nodes)
        return ops.nearest(None, None, self.n_nodes, targets, x32=self.x32, y32=self.y32,
                           coord_bound=coord_bound, max_nodes=max_nodes, states=self.states)

    def bytes_scanned_per_query(self):
        return 8 if self.x32 is not None else 16

    def bytes_per_node(self):
        return 40 + (16 if self.x is not None else 0) + (8 if self.x32 is not None else 0) + \
            (4 if self.parent is not None else 0)


class BatchedExtender:
    """nearest -> gather -> extend for Q queries per call."""

    def __init__(self, env, forest, dwa=None, policy=None, eps=0.0, t_max_extend=10.0, t_tree=5.0, rl_chunk=8192,
                 pipe_depth=3):
        self.env = env
        self.pipe_depth = max(2, int(pipe_depth))   # steps in flight of submit_host / collect
        self.rl_chunk = rl_chunk      # policy steering: queries per crlk_extend_rl pass (bounds its workspace)
        self.forest = forest
        self.dwa = dwa
        self.policy = policy
        self.eps = eps
        self.xp = ops.extend_params(t_max_extend, t_tree)
        self.coord_bound = env.coord_bound()
        Q = forest.Q
        d = forest.states.device
        self.from_states = torch.empty(Q, 5, dtype=torch.float64, device=d)
        self.stats = torch.zeros(8, dtype=torch.int64, device=d)      # crlk_extend_dwa work counters (include/crlk.h)
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
        self.extend_targets = None    # extend toward these instead of the samples (cfg5)
        # kernels per step(): nearest (+ its merge pass for split trees) + gather + the extension
        # (DWA: order + extend; policy steering: what the library says each crlk_extend_rl pass launches)
        if policy is None:
            # k_nearest, k_nearest_final, k_gather_rows, k_lpt_order (batches larger than the persistent grid), k_extend_dwa
            self.launches_per_step = 5 if Q > 296 else 4
        else:
            import ctypes as C
            from .._lib import lib
            n = 3                                                  # k_nearest, k_nearest_final, k_gather_rows
            for q0 in range(0, Q, rl_chunk):
                n += int(lib().crlk_extend_rl_launches(policy.actor.handle.h, C.byref(self.xp), min(Q, q0 + rl_chunk) - q0, 0))
            self.launches_per_step = n

    def step(self, samples, events=None):
        """samples: float64 CUDA [Q, 5].  Returns the extension outputs (CUDA tensors)."""
        f = self.forest
        ev = events

        def mark(i):
            if ev is not None:
                ev[i].record()
        mark(0)
        idx, dist, tie = f.nearest(samples, self.coord_bound, max_nodes=self.max_nodes)
        mark(1)
        ops.gather_rows(f.states, idx, out=self.from_states)
        mark(2)
        if self.policy is None:
            self.out = ops.extend_dwa(self.env.handle, self.dwa.params(), self.xp, self.from_states, samples,
                                      out=self.out, stats=self.stats)
        else:
            tg = samples if self.extend_targets is None else self.extend_targets
            self.out = self._extend_rl_chunked(tg)
        mark(3)
        self.out["nearest_idx"], self.out["nearest_dist"] = idx, dist
        return self.out

    def _extend_rl_chunked(self, targets):
        """crlk_extend_rl over the batch in passes of rl_chunk queries (outputs written in place)."""
        Q, c = self.forest.Q, self.rl_chunk
        if Q <= c:
            return ops.extend_rl(self.env.handle, self.policy.actor.handle, self.xp, self.from_states, targets,
                                 None, self.eps, out=self.out)
        out = self.out if self.out is not None else ops.extend_rl_outputs(self.xp, Q, self.from_states.device)
        for q0 in range(0, Q, c):
            q1 = min(Q, q0 + c)
            view = {k: (v[q0:q1] if torch.is_tensor(v) else v) for k, v in out.items() if not k.startswith("_")}
            r = ops.extend_rl(self.env.handle, self.policy.actor.handle, self.xp, self.from_states[q0:q1],
                              targets[q0:q1], None, self.eps, out=view, workspace=out.get("_ws"))
            out["_ws"] = r["_ws"]
        return out

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

    # ---- pipelined host-buffer entry points ------------------------------------------
    def _pipe_init(self):
        d = self.forest.states.device
        self._pipe = []
        for _ in range(self.pipe_depth):
            self._pipe.append(dict(
                h_in=torch.empty_like(self._h_samples).pin_memory(), d_in=torch.empty_like(self._d_samples),
                h_out={k: torch.empty_like(v).pin_memory() for k, v in self._h_out.items()},
                out=None, from_states=torch.empty_like(self.from_states),
                stream=torch.cuda.Stream(device=d),
                done=torch.cuda.Event(), copied=torch.cuda.Event(), busy=False))
        self._copy_stream = torch.cuda.Stream(device=d)
        self._slot = 0

    def submit_host(self, samples_np):
        """Asynchronous form of step_host: stages the samples (pinned) and enqueues H2D, the step and the
        D2H of its results; returns a ticket for collect().  pipe_depth steps may be in flight, each on its
        own stream: the D2H of one step overlaps the kernels of the next, and -- a batch of extensions ends
        when its longest one does -- the next step's extensions fill the SMs that the tail of the
        previous step leaves idle.  Steps are independent (a step reads the forest, it does not grow
        it); every step still moves its own inputs and results."""
        if not hasattr(self, "_pipe"):
            self._pipe_init()
        i = self._slot
        self._slot = (self._slot + 1) % self.pipe_depth
        s = self._pipe[i]
        if s["busy"]:
            s["copied"].synchronize()              # the slot's previous results must have been collected / copied
        s["h_in"].numpy()[...] = samples_np
        cur = torch.cuda.current_stream()
        s["stream"].wait_stream(cur)               # whatever the caller enqueued so far (forest updates) comes first
        with torch.cuda.stream(s["stream"]):
            s["d_in"].copy_(s["h_in"], non_blocking=True)
            if self.policy is not None and getattr(self, "_last_done", None) is not None:
                # policy steering: the kernels of a step fill the GPU (persistent GEMM grid), two steps side by
                # side only slow each other down -- the next step's kernels wait for the previous step's kernels,
                # not for its D2H, which is what the pipeline hides here
                s["stream"].wait_event(self._last_done)
            keep_out, keep_from = self.out, self.from_states
            self.out, self.from_states = s["out"], s["from_states"]
            out = self.step(s["d_in"])
            s["out"] = out
            self.out, self.from_states = keep_out, keep_from
            s["done"].record()
            self._last_done = s["done"]
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(s["done"])
            for k, h in s["h_out"].items():
                h.copy_(out[k], non_blocking=True)
            s["copied"].record()
        s["busy"] = True
        return i

    def submit(self, samples_dev, flush=None):
        """Device-resident form of submit_host: the step runs on the slot's stream (pipe_depth steps in
        flight), inputs and results stay in HBM; wait() returns the slot's output tensors.  `flush`
        (a CUDA tensor, optional) is overwritten on the slot's stream before the step (bench.py's L2 flush)."""
        if not hasattr(self, "_pipe"):
            self._pipe_init()
        i = self._slot
        self._slot = (self._slot + 1) % self.pipe_depth
        s = self._pipe[i]
        cur = torch.cuda.current_stream()
        s["stream"].wait_stream(cur)
        with torch.cuda.stream(s["stream"]):
            if flush is not None:
                flush.fill_(i)
            keep_out, keep_from = self.out, self.from_states
            self.out, self.from_states = s["out"], s["from_states"]
            s["out"] = self.step(samples_dev)
            self.out, self.from_states = keep_out, keep_from
            s["done"].record()
        return i

    def wait(self, ticket):
        """Makes the current stream wait for a submit() step and returns its output tensors."""
        s = self._pipe[ticket]
        torch.cuda.current_stream().wait_event(s["done"])
        return s["out"]

    def collect(self, ticket):
        """Results of a submit_host() call as numpy views of pinned memory (valid until the slot is reused
        pipe_depth submissions later)."""
        s = self._pipe[ticket]
        s["copied"].synchronize()
        s["busy"] = False
        return {k: h.numpy() for k, h in s["h_out"].items()}

    def h2d_bytes(self):
        return self._h_samples.numel() * 8

    def d2h_bytes(self):
        return sum(h.numel() * h.element_size() for h in self._h_out.values())


class BatchedRRT:
    """Q independent RRT.planning() runs (rrt.py:51-89) advanced in lock step,
    trees resident in HBM.  One round = one sample per unfinished query:
    validity check + nearest node + acceptance (rrt.py:64-69), one extension
    toward the sample (rrt.py:72), node append + goal test (rrt.py:74-77), and
    the steer-to-goal retry (rrt.py:78-83).  Sampling stays on the host as
    pre-drawn per-query streams (the reference's RNG order cannot be reproduced
    across queries); every query consumes its own stream in order, so its tree
    is the one the sequential planner would build from the same stream.

    steering: 'dwa' (RRT) or 'rl' (RRT_RL with `policy`, optional noise stream).
    parent selection: nearest node, or -- with `estimator` (the pair returned by load_estimator) --
    RRT_RL_Estimator.choose_parent (rrt_rl_estimator.py:26-79) for both the sample and the retry.
    """

    def __init__(self, env, starts, goals, max_iter=500, dwa=None, policy=None, eps=0.0,
                 keep_paths=True, node_capacity=None, with_f32=True, estimator=None):
        import ctypes as C
        from .._lib import ExtendView, ForestView
        self.env = env
        self.Q = Q = len(starts)
        self.max_iter = max_iter
        self.policy = policy
        self.eps = eps
        self.estimator = estimator[0].pair if isinstance(estimator, (tuple, list)) else \
            (estimator.pair if estimator is not None and hasattr(estimator, "pair") else estimator)
        if dwa is None and policy is None:
            from ..policy.dwa import DWA
            dwa = DWA(env)
            dwa.set_dwa(dt=0.2, v_res=0.02)          # rrt.py:39-40
        self.dwa = dwa
        self.xp = ops.extend_params()
        self.per = self.xp.n_max_extend // self.xp.n_tree + 1
        d = ops.dev()
        cap = node_capacity or (2 * (self.per - 1) * max_iter + 1)
        self.forest = f = Forest(Q, cap, d, with_f32=with_f32)
        self.starts = ops.as_dev(starts).reshape(Q, 5)
        self.goals = ops.as_dev(goals).reshape(Q, 5)
        v, _, _ = ops.valid_state(env.handle, torch.cat([self.starts, self.goals]))
        assert bool(v.all().item()), "The start or goal states are not valid"      # rrt.py:93
        f.fill(self.starts.reshape(Q, 1, 5))
        self.keep_paths = keep_paths
        self.pool_cap = self.xp.n_max_extend * 2 * max_iter if keep_paths else 0
        if keep_paths:
            self.pool = torch.zeros(Q, self.pool_cap, 5, dtype=torch.float64, device=d)
            self.seg_off = torch.zeros(Q, cap, dtype=torch.int32, device=d)
            self.seg_len = torch.zeros(Q, cap, dtype=torch.int32, device=d)
            self.seg_flag = torch.zeros(Q, cap, dtype=torch.uint8, device=d)
            self.pool_used = torch.zeros(Q, dtype=torch.int32, device=d)
        i32 = dict(dtype=torch.int32, device=d)
        u8 = dict(dtype=torch.uint8, device=d)
        self.iters = torch.zeros(Q, **i32)
        self.done = torch.zeros(Q, **u8)
        self.reached = torch.zeros(Q, **u8)
        self.active = torch.zeros(Q, **u8)
        self.retry_active = torch.zeros(Q, **u8)
        self.retry_from = torch.zeros(Q, 5, dtype=torch.float64, device=d)
        self.retry_parent = torch.zeros(Q, **i32)
        self.overflow = torch.zeros(1, **i32)
        self.cursor = torch.zeros(Q, **i32)
        self.sample = torch.zeros(Q, 5, dtype=torch.float64, device=d)
        self.from_states = torch.zeros(Q, 5, dtype=torch.float64, device=d)
        self.stats = torch.zeros(8, dtype=torch.int64, device=d)
        self.out = [None, None]
        self.rounds = 0
        self.coord_bound = env.coord_bound()
        self.first_has_start = 0 if policy is None else 1
        fv = ForestView(f.x.data_ptr(), f.y.data_ptr(), f.x32.data_ptr() if f.x32 is not None else None,
                        f.y32.data_ptr() if f.y32 is not None else None, f.states.data_ptr(),
                        f.parent.data_ptr(), f.n_nodes.data_ptr(), cap,
                        self.pool.data_ptr() if keep_paths else None,
                        self.seg_off.data_ptr() if keep_paths else None,
                        self.seg_len.data_ptr() if keep_paths else None,
                        self.seg_flag.data_ptr() if keep_paths else None,
                        self.pool_used.data_ptr() if keep_paths else None, self.pool_cap)
        self._fv = fv
        self._C = C
        self._ExtendView = ExtendView
        self.noise = None             # float32 [Q, rows, 2] standard-normal stream per query (eps != 0)
        self.noise_cursor = torch.zeros(Q, dtype=torch.int32, device=d)

    def reset(self):
        """Back to the single-node trees (start states), all counters cleared."""
        f = self.forest
        f.fill(self.starts.reshape(self.Q, 1, 5))
        f.parent.fill_(-1)
        for t in (self.iters, self.done, self.reached, self.active, self.retry_active, self.overflow,
                  self.cursor, self.stats, self.noise_cursor):
            t.zero_()
        if self.keep_paths:
            self.pool_used.zero_()
        self.rounds = 0

    # ---- one extension pass ------------------------------------------------
    def _extend(self, slot, from_states, targets, active):
        if self.policy is None:
            self.out[slot] = ops.extend_dwa(self.env.handle, self.dwa.params(), self.xp, from_states, targets,
                                            active=active, out=self.out[slot], stats=self.stats)
        else:
            self.out[slot] = ops.extend_rl(self.env.handle, self.policy.actor.handle, self.xp, from_states,
                                           targets, self.noise, self.eps, active=active,
                                           noise_cursor=self.noise_cursor if self.noise is not None else None)
        return self.out[slot]

    def _append(self, out, parent_idx, active, primary):
        from .._lib import check, lib
        C = self._C
        ev = self._ExtendView(out["path"].data_ptr(), out["n_steps"].data_ptr(), out["status"].data_ptr(),
                              out["node_step"].data_ptr(), out["n_nodes"].data_ptr())
        p = lambda t: C.c_void_p(t.data_ptr())
        check(lib().crlk_rrt_append(C.byref(self._fv), C.byref(ev), p(parent_idx), p(active), p(self.goals),
                                    self.Q, self.xp.n_max_extend, self.xp.n_tree, self.first_has_start,
                                    int(primary), self.max_iter, 1.0, 5.0, p(self.iters), p(self.done),
                                    p(self.reached), p(self.retry_active), p(self.retry_from),
                                    p(self.retry_parent), p(self.overflow), ops._stream()))

    def round(self, samples):
        """Advance every unfinished query by one sample (samples: CUDA float64 [Q, 5])."""
        from .._lib import check, lib
        C = self._C
        f = self.forest
        p = lambda t: C.c_void_p(t.data_ptr())
        valid, _, _ = ops.valid_state(self.env.handle, samples)                       # rrt.py:66
        est = self.estimator
        # every round appends at most 2 * (per - 1) nodes per tree: a host-side bound of the tree sizes
        bound = min(f.cap, 1 + 2 * (self.per - 1) * (self.rounds + 1))
        if est is None:
            idx, dist, _ = ops.nearest(f.x, f.y, f.n_nodes, samples, x32=f.x32, y32=f.y32,
                                       coord_bound=self.coord_bound)                   # rrt.py:68
        else:
            idx, dist = ops.estimator_choose_parent(self.env.handle, est.handle, f.states, f.n_nodes, samples,
                                                    max_nodes=bound)                   # rrt_rl_estimator.py:26
        check(lib().crlk_rrt_select(p(valid), p(dist), self.Q, self.max_iter, 1.0, 20.0, p(self.done),
                                    p(self.iters), p(self.active), ops._stream()))      # rrt.py:69
        ops.gather_rows(f.states, idx, out=self.from_states)                          # rrt.py:70
        out = self._extend(0, self.from_states, samples, self.active)                 # rrt.py:72
        if est is None:
            self._append(out, idx, self.active, True)                                 # rrt.py:74-79
        else:
            n_before = f.n_nodes.clone()
            self._append(out, idx, self.active, 2)                                    # rrt.py:74-77
            ridx, rcost = ops.estimator_choose_parent(self.env.handle, est.handle, f.states, f.n_nodes,
                                                      self.goals, max_nodes=min(f.cap, bound + 2 * (self.per - 1)),
                                                      first_node=n_before)             # rrt.py:78
            check(lib().crlk_rrt_retry_choice(C.byref(self._fv), p(n_before), p(ridx), p(rcost), p(self.active),
                                              self.Q, 5.0, self.max_iter, p(self.iters), p(self.done),
                                              p(self.retry_active), p(self.retry_from), p(self.retry_parent),
                                              ops._stream()))                          # rrt.py:79
        out2 = self._extend(1, self.retry_from, self.goals, self.retry_active)        # rrt.py:80
        self._append(out2, self.retry_parent, self.retry_active, False)               # rrt.py:81-83
        self.rounds += 1

    def plan(self, streams, check_every=8, max_rounds=None, noise=None):
        """streams: float64 [Q, S, 5] pre-drawn samples per query (host or device).
        noise: float32 [Q, rows, 2] exploration draws per query (RL steering with eps != 0).
        Runs rounds until every query is done or its stream is exhausted."""
        streams = ops.as_dev(streams)
        if noise is not None:
            self.noise = ops.as_dev(noise, torch.float32).contiguous()
        S = streams.shape[1]
        max_rounds = min(S, max_rounds or S)
        step = torch.zeros(self.Q, dtype=torch.int32, device=streams.device)
        for r in range(max_rounds):
            step.fill_(r)
            ops.gather_rows(streams, step, out=self.sample)
            self.round(self.sample)
            if (r + 1) % check_every == 0 and bool(self.done.all().item()):
                break
        torch.cuda.synchronize()
        if int(self.overflow.item()):
            raise RuntimeError(f"{int(self.overflow.item())} trees ran out of node/path capacity")
        if self.noise is not None and self.eps != 0.0:
            # one draw per executed policy step (net.py:147-150); the kernel treats draws past the end of
            # a query's stream as zero, which would silently change the exploration: refuse that
            used = int(self.noise_cursor.max().item())
            if used > self.noise.shape[1]:
                raise RuntimeError(f"a query consumed {used} exploration draws but its noise stream holds "
                                   f"{self.noise.shape[1]}: pass at least 2 * max_iter * n_max_extend rows")
        self.samples_used = r + 1
        return self.reached.cpu().numpy().astype(bool)

    # ---- results ---------------------------------------------------------------
    def tree(self, q):
        """(states [n,5], parents [n]) of query q as numpy arrays."""
        n = int(self.forest.n_nodes[q].item())
        return self.forest.states[q, :n].cpu().numpy(), self.forest.parent[q, :n].cpu().numpy()

    def final_course(self, q, goal_ind=-1):
        """generate_final_course (rrt.py:137-144) for query q."""
        assert self.keep_paths, "paths were not kept"
        states, parents = self.tree(q)
        n = len(states)
        off = self.seg_off[q, :n].cpu().numpy()
        ln = self.seg_len[q, :n].cpu().numpy()
        flag = self.seg_flag[q, :n].cpu().numpy()
        used = int(self.pool_used[q].item())
        pool = self.pool[q, :used].cpu().numpy()
        node = goal_ind % n
        path = []
        while parents[node] >= 0:
            seg = [pool[off[node] + i] for i in range(ln[node])]
            full = ([states[parents[node]]] if flag[node] else []) + seg     # Node.path
            path = full[1:] + path
            node = parents[node]
        return [states[node]] + path


class FusedRRT(BatchedRRT):
    """Q independent RRT.planning() runs (rrt.py:51-89, DWA steering) in ONE kernel
    launch (crlk_plan_dwa): a thread block owns a query from its first sample to
    its last node, persistent blocks pull queries from a work counter, so no query
    waits for another (BatchedRRT advances all queries in lock step, one sample
    per round, with five launches per round).  Same forest, same pre-drawn sample
    streams, same trees and paths as BatchedRRT / the sequential planner."""

    def __init__(self, env, starts, goals, max_iter=500, dwa=None, keep_paths=True, node_capacity=None,
                 with_f32=False):
        super().__init__(env, starts, goals, max_iter=max_iter, dwa=dwa, policy=None, keep_paths=keep_paths,
                         node_capacity=node_capacity, with_f32=with_f32)
        d = self.starts.device
        self.flags = torch.zeros(self.Q, dtype=torch.uint8, device=d)
        # scheduling hint: queries whose goal is far from the start tend to need the most iterations;
        # handing those out first shortens the tail of the launch (results do not depend on the order)
        gap = (self.starts[:, :2] - self.goals[:, :2]).norm(dim=1)
        self.order = torch.argsort(gap, descending=True).to(torch.int32).contiguous()

    def plan_params(self, seed=0, goal_sample_rate=5):
        from .._lib import PlanParams
        C = self._C
        sb = self.env.get_bounds()["state_bounds"]
        return PlanParams(self.max_iter, 1.0, 20.0, 1.0, 5.0, int(seed) & 0xFFFFFFFFFFFFFFFF,    # rrt.py:69, :75, :79
                          (C.c_double * 5)(*sb[:, 0]), (C.c_double * 5)(*sb[:, 1]), int(goal_sample_rate))

    def sample_streams(self, n, seed=0, goal_sample_rate=5):
        """The draws plan_async(None, ...) consumes, materialised on the device: float64 [Q, n, 5]."""
        from .._lib import check, lib
        C = self._C
        out = torch.empty(self.Q, n, 5, dtype=torch.float64, device=self.goals.device)
        pp = self.plan_params(seed, goal_sample_rate)
        check(lib().crlk_sample_stream(C.byref(pp), C.c_void_p(self.goals.data_ptr()), self.Q, n,
                                       C.c_void_p(out.data_ptr()), ops._stream()))
        return out

    def plan_async(self, streams, seed=0, max_samples=None, goal_sample_rate=5):
        """Launch the planning kernel on the current stream.  streams: CUDA float64 [Q, S, 5], or None:
        the kernel then draws its own samples (RRT.sample with a counter-based generator keyed by
        `seed`, at most max_samples per query; workloads.device_sample_stream is the host mirror)."""
        from .._lib import check, lib
        C = self._C
        p = lambda t: C.c_void_p(t.data_ptr())
        pp = self.plan_params(seed, goal_sample_rate)
        if streams is None:
            S = int(max_samples or 8 * self.max_iter + 64)
            sptr = None
        else:
            assert streams.is_cuda and streams.dtype == torch.float64 and streams.is_contiguous()
            assert streams.shape[0] == self.Q and streams.shape[2] == 5
            S = streams.shape[1]
            sptr = p(streams)
        self._streams = streams
        check(lib().crlk_plan_dwa(self.env.handle.h, C.byref(self.dwa.params()), C.byref(self.xp), C.byref(pp),
                                  C.byref(self._fv), p(self.goals), sptr, S, self.Q,
                                  p(self.iters), p(self.cursor), p(self.reached), p(self.flags),
                                  p(self.overflow), p(self.stats), p(self.order) if self.order is not None else None,
                                  ops._stream()))

    def plan(self, streams=None, seed=0, max_samples=None, **_):
        if streams is not None:
            streams = ops.as_dev(streams).contiguous()
        self.plan_async(streams, seed=seed, max_samples=max_samples)
        torch.cuda.synchronize()
        if int(self.overflow.item()):
            raise RuntimeError(f"{int(self.overflow.item())} trees ran out of node/path capacity")
        self.done = ((self.reached != 0) | (self.iters >= self.max_iter)).to(torch.uint8)
        self.samples_used = int(self.cursor.max().item())
        return self.reached.cpu().numpy().astype(bool)
