#!/usr/bin/env python
"""bench.py -- RRT extensions/s of the batched extend path on B200.

One "step" = one pass of the hot path over one batch: for each of Q queries, nearest node of its
sample over the query's tree, then one whole extension (<= 50 closed-loop control steps with motion
and collision checks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|cfg4|cfg5]

Main line = BASELINE config 4 (dense 10,004-obstacle map, 1024 concurrent queries per GPU, DWA
steering with a 64 x 64 nominal rollout grid; weak scaling).  With N > 1 (torchrun, one rank per GPU,
NCCL) the line also carries BASELINE config 5 under "cfg5": 65,536 queries sharded over the ranks,
100,000-node trees, batched policy steering, and the final NCCL gather of the per-query results on
rank 0.  Queries are independent: no collective on the data path.

Both arms (--impl ours / reference) run on the SAME data: crl_kino_b200.workloads.cfg4_workload is
generated on the host from fixed seeds (prefix-stable in queries and batches); the reference arm
times the CPU restatement of the reference path (oracle/crlk_oracle.c, all host threads) on a bounded
prefix of every batch.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- work-unit constants -----------------------------------------------
M_FLOP_ROLLOUT = 406.0      # FP ops per DWA rollout as the reference evaluates it
M_FLOP_PAIR = 860.0         # FP ops per (pose, obstacle) pair, brute force
FP64_LANES_PER_SM = 64      # B200: 64 FP64 FMA lanes per SM
# float64 operations k_extend_dwa executes per counted unit (FMA = 2; DESIGN.md section 4 derives them from
# the source and checks the total against ncu's dadd + dmul + 2 dfma count of one launch: within 8 %)
X_FLOP = dict(rollout=99.0, sincos_task=130.0, velocity_chain=100.0, pair_test=400.0, step_fixed=300.0)
WORKLOAD_SEED = 100
POLICY_FLOP = 3600384.0     # 2 * MACs of the 733-1024-512-512-512-2 actor


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "cfg4", "cfg5"],
                    help="auto = cfg4, plus the cfg5 leg when more than one GPU is used")
    ap.add_argument("--queries", type=int, default=0, help="queries per GPU (0 = 1024 for cfg4, 65536 / gpus for cfg5)")
    ap.add_argument("--tree-nodes", type=int, default=1024)
    ap.add_argument("--cells", type=int, default=10000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity legs")
    ap.add_argument("--no-pyref", action="store_true", help="skip the timing of the unmodified Python reference (cfg1)")
    ap.add_argument("--use-clearance", action="store_true",
                    help="enable the per-rollout clearance term (dwa.py:64-69); GPU legs only")
    ap.add_argument("--no-extra", action="store_true", help="skip the planning / nearest / MLP side measurements")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="length of the sustained step loop (0 = skip)")
    ap.add_argument("--pipe-depth", type=int, default=3, help="independent steps in flight of the end-to-end arm")
    ap.add_argument("--plan-iters", type=int, default=100, help="max_iter of the planning-queries/s leg")
    ap.add_argument("--plan-queries", type=int, default=0, help="queries per GPU of the planning leg (0 = --queries)")
    ap.add_argument("--plan-lockstep", action="store_true", help="also time the lock-step planner (BatchedRRT)")
    ap.add_argument("--plan-rl-queries", type=int, default=8192,
                    help="concurrent queries of the learned-steering planning leg (lock-step rounds: the rate grows with the "
                         "batch, 2.7 k queries/s at 1024, 8.9 k at 4096, 16.8 k at 8192)")
    ap.add_argument("--plan-rl-iters", type=int, default=6, help="max_iter of the learned-steering planning leg")
    ap.add_argument("--cfg5-steps", type=int, default=4)
    ap.add_argument("--cfg5-warmup", type=int, default=2)
    ap.add_argument("--cfg5-total", type=int, default=65536, help="queries of the cfg5 leg over all GPUs")
    ap.add_argument("--cfg5-nodes", type=int, default=100000)
    ap.add_argument("--pyref", nargs=2, type=int, metavar=("SEED", "MAX_ITER"), help=argparse.SUPPRESS)
    return ap.parse_args()


def dist_info():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def load_ncu(name):
    """Summary of an `ncu --set full` capture of the same command, written by tools/ncu_summary.py
    from the .ncu-rep (profiles/<name>.json); None when no capture is committed."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        d["file"] = "profiles/" + name
        return d
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons under load.  The sampler is started ahead of the load
    (nvidia-smi needs a few hundred ms to print its first row); every row is stamped on arrival and
    stop() keeps the rows that fall inside the given windows of time.perf_counter()."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=3.0):
        """Block until nvidia-smi has printed its first row (so the load that follows is sampled)."""
        t0 = time.perf_counter()
        while self.p is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, windows=None):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            # a row printed up to one sampling period after a window still describes it
            if windows is not None and not any(a <= t <= b + 0.03 for a, b in windows):
                continue
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ the shared cfg4 workload

def cfg4_config(args, n_obstacles):
    """The `config` object of BOTH arms (identical for identical flags)."""
    return {"workload": "cfg4: dense random-obstacle map, 1024 concurrent queries, DWA steering, 64 x 64 nominal rollout grid",
            "obstacles": int(n_obstacles), "queries_per_gpu": int(args.queries), "tree_nodes": int(args.tree_nodes),
            "rollouts_per_control_nominal": 4096, "use_clearance": bool(args.use_clearance),
            "generator": f"crl_kino_b200.workloads.cfg4_workload(seed={WORKLOAD_SEED} + rank)",
            "l2": "flushed between timed iterations (256 MiB write)"}


def workload_fingerprint(w, W):
    """sha256 over the first 64 queries of timed batch 0 and their trees (both arms can compute it)."""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(w["batches"][W, :64]).tobytes())
    h.update(np.ascontiguousarray(w["trees"][:64]).tobytes())
    return h.hexdigest()[:16]


def timed(fn, n, warm=2):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# ------------------------------------------------------------------ CPU legs (oracle = the checker / baseline)

def cpu_sample_run(rects, tree_states, samples, n, threads, want_ctrl=False):
    """The oracle's C restatement of nearest + DWA extension on the first n queries."""
    from crl_kino_b200.workloads import DENSE_DWA
    from oracle import oracle as orc
    o = orc.Env()
    o.set_obs(rects)
    d = orc.Dwa(o, **DENSE_DWA)
    xs = np.ascontiguousarray(tree_states[:n, :, 0])
    ys = np.ascontiguousarray(tree_states[:n, :, 1])
    t0 = time.perf_counter()
    idx, dist = orc.nearest_batch(xs, ys, np.full(n, xs.shape[1], dtype=np.int32), samples[:n], threads=threads)
    froms = tree_states[np.arange(n), idx]
    r = d.extend_batch(froms, samples[:n], threads=threads, want_ctrl=want_ctrl)
    dt = time.perf_counter() - t0
    r["nearest_idx"], r["nearest_dist"] = idx, dist
    return n / dt, dt, r


def parity_block(gpu, ref, n):
    """GPU results of timed batch 0 against the oracle's on the same n queries.  identical = same
    nearest node, step count, status, emitted nodes, chosen rollout at every control step, and a path
    within 1e-9 relative; near_boundary / near_tie = the kernel's own report bits (flags bit 0 / 1);
    diverged = not identical AND not reported."""
    fl = gpu["flags"][:n]
    same = np.ones(n, dtype=bool)
    bitwise = 0
    worst = 0.0
    for q in range(n):
        ok = (gpu["nearest_idx"][q] == ref["nearest_idx"][q] and gpu["n_steps"][q] == ref["n_steps"][q] and
              gpu["status"][q] == ref["status"][q] and gpu["n_nodes"][q] == ref["n_nodes"][q])
        if ok:
            m = int(ref["n_steps"][q])
            ok = np.array_equal(gpu["ctrl_idx"][q, :m], ref["ctrl_idx"][q, :m])
            a, b = gpu["path"][q, :m], ref["path"][q, :m]
            if ok and m:
                err = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3)))
                worst = max(worst, err)
                ok = err <= 1e-9
                bitwise += int(np.array_equal(a, b))
            elif ok:
                bitwise += 1
        same[q] = ok
    return {"compared": int(n), "identical": int(same.sum()), "bitwise_identical_paths": int(bitwise),
            "near_boundary": int(((fl & 1) != 0).sum()), "near_tie": int(((fl & 2) != 0).sum()),
            "diverged": int((~same & (fl == 0)).sum()), "diverged_reported": int((~same & (fl != 0)).sum()),
            "worst_path_rel_err_of_identical": worst,
            "contract": "indices / flags bit-exact, paths <= 1e-4 relative (checked at 1e-9 and bitwise); reported separately: "
                        "states within 1e-6 of an obstacle boundary (near_boundary) and argmin decisions that one misrounded "
                        "last bit of the C library's atan2 could change (near_tie, include/crlk.h crlk_dwa_control)"}


def python_reference_cfg1(max_iter=30, seeds=(0, 1)):
    """The UNMODIFIED Python reference (baseline/_ref, installed from /root/reference with pip
    --target) on BASELINE config 1: examples/test.py --case rrt on test_env1, one process per seed,
    max_iter bounded so that the leg takes ~30 s.  Extensions = RRT.steer calls (rrt.py:72,80)."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "crl_kino")):
        return {"unavailable": "baseline/_ref is not installed (python -m pip install --no-index --no-deps "
                               "--target baseline/_ref /root/reference)"}
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--pyref", str(s), str(max_iter)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for s in seeds]
    runs = []
    for p in procs:
        out, err = p.communicate()
        try:
            runs.append(json.loads(out.strip().splitlines()[-1]))
        except Exception:
            return {"unavailable": "reference run failed: " + (err.strip().splitlines()[-1] if err.strip() else "no output")}
    ext = sum(r["extensions"] for r in runs)
    sec = max(r["seconds"] for r in runs)
    return {"ext_per_s": ext / sec, "ext_per_s_per_core": float(np.mean([r["extensions"] / r["seconds"] for r in runs])),
            "seconds": sec, "extensions": ext, "control_steps": sum(r["control_steps"] for r in runs),
            "cores": len(runs), "max_iter": max_iter, "seeds": list(seeds),
            "what": "crl_kino.planner.rrt.RRT.planning() of the unmodified reference (rrt.py:51-89), test_env1, start "
                    "[-5,-15,0,0,0] -> goal [10,10,0,0,0], one process per seed"}


def pyref_child(seed, max_iter):
    """Child process of python_reference_cfg1: imports the reference, never this package's kernels."""
    import contextlib
    import io
    import random
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden", "_stubs"))       # matplotlib / gym / transforms3d import stubs
    from crl_kino.env import DifferentialDriveEnv
    from crl_kino.planner.rrt import RRT
    from crl_kino.utils.rigid import RectObs
    rects = np.load(os.path.join(ROOT, "data", "fixtures.npz"))["map_test_env1"]
    env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi)
    env.set_obs([RectObs(r.reshape(2, 2)) for r in rects])
    p = RRT(env, max_iter=max_iter)
    p.set_start_and_goal(np.array([-5, -15, 0, 0, 0.]), np.array([10, 10, 0, 0, 0.]))
    n = {"ext": 0, "ctl": 0}
    steer0, ctl0 = p.steer, p.dwa.control

    def steer(*a, **k):
        n["ext"] += 1
        return steer0(*a, **k)

    def ctl(x, g):
        n["ctl"] += 1
        return ctl0(x, g)
    p.steer, p.dwa.control = steer, ctl
    random.seed(seed)
    np.random.seed(seed)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        p.planning()
    print(json.dumps({"seed": seed, "extensions": n["ext"], "control_steps": n["ctl"], "seconds": time.perf_counter() - t0}))


def _cpu_plan_one(a):
    rects, start, goal, stream, iters = a
    from crl_kino_b200.workloads import DENSE_DWA
    from oracle import oracle as orc
    o = orc.Env()
    o.set_obs(rects)
    p = orc.PlannerPort(o, max_iter=iters, dwa_kwargs=DENSE_DWA)
    p.set_start_and_goal(start, goal)
    try:
        p.planning(iter(stream))
    except StopIteration:                      # ran past the mirrored draws: a decision went the other way
        return p.n_extend_calls, False, -1
    return p.n_extend_calls, bool(p.reach_exactly), len(p.node_list)


def cpu_planning(rects, inputs, iters, n, cores):
    import multiprocessing as mp
    starts, goals, streams = inputs
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_plan_one, [(rects, starts[q], goals[q], streams[q], iters) for q in range(n)])
    dt = time.perf_counter() - t0
    return {"queries_per_sec": n / dt, "queries": n, "seconds": dt, "cores": cores, "kind": "port",
            "extensions": int(sum(r[0] for r in res)), "reached_frac": float(np.mean([r[1] for r in res])),
            "_nodes": [r[2] for r in res], "_reached": [r[1] for r in res]}


# ------------------------------------------------------------------ side measurements (rank 0, N = 1 figures)

def planning_leg(args, env, dwa, rank):
    """Planning queries/s: Q full RRT.planning runs (max_iter = --plan-iters) in one launch of
    crlk_plan_dwa (FusedRRT, one CTA per query).  The headline figure lets the kernel draw RRT.sample
    itself (counter-based generator): every query runs until it reaches its goal or max_iter, none
    stops on an exhausted stream.  `from_streams` repeats it from pre-drawn host streams (the form the
    parity tests use), with the H2D of the streams timed in its e2e."""
    import torch
    from crl_kino_b200.planner.batched import BatchedRRT, FusedRRT
    from crl_kino_b200.workloads import device_sample_stream, query_pairs
    Q = args.plan_queries or args.queries

    def clearance(c):
        return env.valid_state_batch(c, want_clearance=True)[1].cpu().numpy()
    starts, goals = query_pairs(clearance, Q, seed=7 + rank)
    sb = env.get_bounds()["state_bounds"]
    S = 64 * args.plan_iters                       # far more draws than any query consumes
    seed = 300 + rank
    out = {"queries": Q, "max_iter": args.plan_iters}
    fp = FusedRRT(env, starts, goals, max_iter=args.plan_iters, dwa=dwa, keep_paths=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for rep in range(3):                           # first pass = warm-up, best of the next two
        fp.reset()
        torch.cuda.synchronize()
        e0.record()
        fp.plan_async(None, seed=seed, max_samples=S)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if rep > 0:
            best = ms if best is None else min(best, ms)
    st = fp.stats.cpu().numpy()
    reached = fp.reached.cpu().numpy().astype(bool)
    cursor = fp.cursor.cpu().numpy()
    iters = fp.iters.cpu().numpy()
    out.update(seconds=best * 1e-3, reached_frac=float(reached.mean()), extensions=int(st[0]), control_steps=int(st[1]),
               mean_nodes=float(fp.forest.n_nodes.double().mean().item()), mean_samples=float(cursor.mean()),
               max_samples_consumed=int(cursor.max()),
               stopped_on_sample_limit=int(((cursor >= S) & ~reached & (iters < args.plan_iters)).sum()))
    # end to end: launch + D2H of the per-query results, wall clock
    h_res = [torch.empty(Q, dtype=torch.uint8).pin_memory(), torch.empty(Q, dtype=torch.int32).pin_memory(),
             torch.empty(Q, dtype=torch.int32).pin_memory()]
    h_q = torch.as_tensor(np.concatenate([starts, goals], 1)).pin_memory()
    d_q = torch.empty_like(h_q, device="cuda")
    fp.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d_q.copy_(h_q, non_blocking=True)              # the queries themselves (start, goal) are the inputs
    fp.plan_async(None, seed=seed, max_samples=S)
    for h, t in zip(h_res, (fp.reached, fp.iters, fp.forest.n_nodes)):
        h.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    out["e2e_seconds"] = time.perf_counter() - t0
    out["h2d_bytes"] = h_q.numel() * 8
    out["d2h_bytes"] = sum(h.numel() * h.element_size() for h in h_res)
    # the host mirror of the draws, for the CPU leg (only the first queries / draws it needs)
    n_cpu = min(Q, 64)
    need = int(cursor[:n_cpu].max()) + 64
    out["inputs"] = (starts[:n_cpu], goals[:n_cpu], device_sample_stream(sb, goals[:n_cpu], need, seed))
    out["gpu_first"] = dict(reached=reached[:n_cpu].copy(), nodes=fp.forest.n_nodes.cpu().numpy()[:n_cpu].copy())
    # pre-drawn host streams (the parity tests' form)
    Ss = args.plan_iters * 4 + 32
    rng = np.random.default_rng(300 + rank)
    streams = rng.uniform(sb[:, 0], sb[:, 1], (Q, Ss, 5))
    pick = rng.integers(0, 101, (Q, Ss)) <= 5                      # goal bias of rrt.py:147
    streams[pick] = np.repeat(goals[:, None, :], Ss, 1)[pick]
    h_streams = torch.as_tensor(streams).pin_memory()
    d_in = torch.empty(h_streams.shape, dtype=torch.float64, device="cuda")
    fp.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d_in.copy_(h_streams, non_blocking=True)
    fp.plan_async(d_in)
    for h, t in zip(h_res, (fp.reached, fp.iters, fp.forest.n_nodes)):
        h.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    out["streams_e2e_seconds"] = time.perf_counter() - t0
    out["streams_h2d_bytes"] = h_streams.numel() * 8
    out["streams_exhausted"] = int(((fp.cursor.cpu().numpy() >= Ss) & (h_res[0].numpy() == 0) &
                                    (h_res[1].numpy() < args.plan_iters)).sum())
    if args.plan_lockstep:
        nodes_from_streams = fp.forest.n_nodes.clone()
        r_streams = h_res[0].numpy().astype(bool).copy()
        p = BatchedRRT(env, starts, goals, max_iter=args.plan_iters, dwa=dwa, keep_paths=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r2 = p.plan(d_in, check_every=16)
        out["lockstep_seconds"] = time.perf_counter() - t0
        out["lockstep_rounds"] = p.samples_used
        out["lockstep_identical"] = bool(np.array_equal(r2, r_streams) and torch.equal(p.forest.n_nodes, nodes_from_streams))
    return out


def _one_thread():
    """Pool initializer: one BLAS / OpenMP thread per worker process (the pool already uses every core)."""
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    try:
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass


def _cpu_plan_rl_one(a):
    rects, start, goal, stream, iters, aw, ab, est, cls = a
    from oracle import oracle as orc
    o = orc.Env()
    o.set_obs(rects)
    p = orc.PlannerPort(o, steering="rl", actor=(aw, ab), estimator=est, classifier=cls, max_iter=iters, eps=0.0)
    p.set_start_and_goal(start, goal)
    try:
        p.planning(iter(stream))
    except StopIteration:
        # the pre-drawn stream ran out inside the rejection loop (rrt.py:64-69 has no cap; the lock-step
        # planner stops such a query with the tree it has): same stop, compare the tree built so far
        return p.n_extend_calls, False, len(p.node_list), True
    return p.n_extend_calls, bool(p.reach_exactly), len(p.node_list), False


def planning_rl_leg(args, env, rects, no_cpu):
    """Planning queries/s with the learned steering function (examples/benchmark.py runs RRT_RL_Estimator
    only, benchmark.py:48-51): Q queries advanced in lock step by BatchedRRT -- per round one nearest-node
    (RRT_RL) or estimator choose_parent (RRT_RL_Estimator) pass, one batched policy extension toward the
    sample and one toward the goal (rrt.py:72-83) -- from pre-drawn host sample streams, wall clock.  The
    CPU port (oracle.PlannerPort, numpy actor / CNN) runs the first queries from the same streams."""
    import multiprocessing as mp
    import torch
    from crl_kino_b200.env import DifferentialDriveGym
    from crl_kino_b200.estimator import load_estimator
    from crl_kino_b200.planner.batched import BatchedRRT
    from crl_kino_b200.policy.rl_policy import load_policy
    from crl_kino_b200.workloads import query_pairs
    Q, iters = args.plan_rl_queries, args.plan_rl_iters
    pol = load_policy(DifferentialDriveGym(), [1024, 512, 512, 512], None, seed=0)
    w = np.load(os.path.join(ROOT, "data", "estimator_weights.npz"))
    est_sd = {k[4:]: w[k] for k in w.files if k.startswith("est.")}
    cls_sd = {k[4:]: w[k] for k in w.files if k.startswith("cls.")}
    est = load_estimator(est_sd, cls_sd)

    def clearance(c):
        return env.valid_state_batch(c, want_clearance=True)[1].cpu().numpy()
    starts, goals = query_pairs(clearance, Q, seed=11)
    sb = env.get_bounds()["state_bounds"]
    S = iters * 6 + 32
    rng = np.random.default_rng(411)
    streams = rng.uniform(sb[:, 0], sb[:, 1], (Q, S, 5))
    pick = rng.integers(0, 101, (Q, S)) <= 5                      # goal bias of rrt.py:147
    streams[pick] = np.repeat(goals[:, None, :], S, 1)[pick]
    h_streams = torch.as_tensor(streams).pin_memory()
    out = {"queries": Q, "max_iter": iters, "unit": "queries/s",
           "policy": "seeded random-init actor 733-1024-512-512-512-2, eps 0 (the trained policy.pth is missing upstream)",
           "estimator": "data/estimator_weights.npz (the reference's dist_est / classifier state dicts)"}
    res = {}
    for name, e in (("rrt_rl", None), ("rrt_rl_estimator", est)):
        best = None
        for rep in range(2):                                    # first pass = warm-up
            p = BatchedRRT(env, starts, goals, max_iter=iters, policy=pol, eps=0.0, keep_paths=False, estimator=e)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            d = h_streams.cuda(non_blocking=True)                # the streams are the step's host inputs
            reached = p.plan(d, check_every=8)
            nodes = p.forest.n_nodes.cpu().numpy()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        res[name] = dict(reached=reached, nodes=nodes)
        out[name] = {"queries_per_sec": Q / best, "seconds": best, "rounds": int(p.samples_used),
                     "reached_frac": float(np.mean(reached)), "mean_nodes": float(nodes.mean()),
                     "h2d_bytes": int(h_streams.numel() * 8), "d2h_bytes": int(nodes.nbytes + Q),
                     "planner": "BatchedRRT(policy" + (", estimator)" if e is not None else ")") + ": lock-step rounds, wall clock incl. H2D / D2H"}
    if not no_cpu:
        from oracle import oracle as orc
        cores = orc.cpu_count()
        n = min(Q, cores, 16)
        aw, ab = pol.actor.weights, pol.actor.biases
        for name, es, cs in (("rrt_rl", None, None), ("rrt_rl_estimator", est_sd, cls_sd)):
            t0 = time.perf_counter()
            with mp.get_context("fork").Pool(n, initializer=_one_thread) as pool:
                r = pool.map(_cpu_plan_rl_one, [(rects, starts[q], goals[q], streams[q], iters, aw, ab, es, cs) for q in range(n)])
            dt = time.perf_counter() - t0
            g = res[name]
            out[name]["cpu_baseline"] = {"queries_per_sec": n / dt, "queries": n, "seconds": dt, "cores": cores, "kind": "port",
                                         "extensions": int(sum(x[0] for x in r)),
                                         "stopped_on_exhausted_stream": int(sum(1 for x in r if x[3])),
                                         "same_reach_and_node_count_as_gpu": int(sum(1 for q in range(n) if r[q][1] == bool(g["reached"][q]) and
                                                                                     r[q][2] == int(g["nodes"][q])))}
    return out


def nearest_roofline(peaks):
    """K3 at cfg5 tree size: 100,000-node trees, float32 SoA scan (8 B per node)."""
    import torch
    from crl_kino_b200 import ops
    Q, N = 1024, 100000
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.rand(Q, N, generator=g, device="cuda", dtype=torch.float64) * 40 - 20)
    y = (torch.rand(Q, N, generator=g, device="cuda", dtype=torch.float64) * 40 - 20)
    x32, y32 = x.float(), y.float()
    n = torch.full((Q,), N, dtype=torch.int32, device="cuda")
    t = (torch.rand(Q, 2, generator=g, device="cuda", dtype=torch.float64) * 40 - 20)
    ms = timed(lambda: ops.nearest(x, y, n, t, x32=x32, y32=y32, coord_bound=22.0), 10)
    algo = 8.0 * Q * N
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = algo / (ms * 1e-3) / 1e9
    ncu = load_ncu("r2_nearest_ncu.json")
    return {"kernel": "k_nearest", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "traffic": ncu.get("dram_bytes") if ncu else None, "ncu": ncu,
            "algorithmic_bytes_per_launch": algo, "ms": ms,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
            "workload": f"{Q} trees x {N} nodes (819 MB of float32 coordinates, > L2), one target each",
            "nodes_per_sec": Q * N / (ms * 1e-3)}


def mlp_roofline(peaks):
    """K4: actor MLP 733-1024-512-512-512-2 at B = 8192 on tcgen05 (bf16 operand split)."""
    import torch
    from crl_kino_b200 import ops
    from crl_kino_b200.env import DifferentialDriveGym
    from crl_kino_b200.policy.rl_policy import load_policy
    B = 8192
    pol = load_policy(DifferentialDriveGym(), [1024, 512, 512, 512], None, seed=0)
    obs = torch.zeros(B, ops.OBS_LD, dtype=torch.float32, device="cuda")
    obs[:, :733] = torch.rand(B, 733, device="cuda")
    h = pol.actor.handle
    ms = timed(lambda: ops.policy_forward(h, obs), 20)
    algo = POLICY_FLOP * B
    peak = float(peaks.get("bf16_tflops", 1590.0))
    ach = algo / (ms * 1e-3) / 1e12
    big = torch.zeros(65536, ops.OBS_LD, dtype=torch.float32, device="cuda")
    big[:, :733] = torch.rand(65536, 733, device="cuda")
    ms_big = timed(lambda: ops.policy_forward(h, big), 5)
    ach_big = POLICY_FLOP * 65536 / (ms_big * 1e-3) / 1e12
    ncu = load_ncu("r2_mlp_chain_ncu.json")
    return {"kernel": "k_mlp_chain (all hidden layers in one dataflow launch) + operand split + head; "
                      "k_gemm_bf16x3_pair<256> per layer at 65,536 rows",
            "bound": "tensor", "achieved": ach, "peak": peak,
            "unit": "TFLOP/s", "frac": ach / peak, "traffic": ncu.get("dram_bytes") if ncu else None, "ms": ms, "ncu": ncu,
            "at_65536_rows": {"achieved": ach_big, "frac": ach_big / peak, "ms": ms_big},
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)" if "bf16_tflops" in peaks else "fallback",
            "note": "algorithmic float32 FLOP (3,600,384 per sample); the kernel issues several bf16 MMAs per "
                    "float32 product (operand split) to keep float32-level accuracy, so the tensor pipe does a "
                    "multiple of this work", "policy_samples_per_sec": B / (ms * 1e-3)}


def sub_rates(env, dwa, peaks):
    """SURVEY 8d sub-rates on the cfg4 map, each kernel alone on resident inputs (CUDA events)."""
    import torch
    from crl_kino_b200 import ops
    from crl_kino_b200.workloads import random_states
    rng = np.random.default_rng(17)
    M = len(env.rects)
    B = 1 << 17
    states = ops.as_dev(random_states(rng, B, moving=True))
    goals = ops.as_dev(random_states(rng, B))
    out = {}
    ms = timed(lambda: ops.valid_state(env.handle, states), 10)
    out["valid_state_checks_per_sec"] = B / (ms * 1e-3)
    out["pose_obstacle_pairs_per_sec_algorithmic"] = B * M / (ms * 1e-3)      # 860 FP ops each, brute force
    ms = timed(lambda: ops.valid_state(env.handle, states, want_clearance=True), 10)
    out["clearances_per_sec"] = B / (ms * 1e-3)
    Bo = 1 << 14
    ms = timed(lambda: ops.local_obs(env.handle, states[:Bo], goals[:Bo], want64=False, want32=True), 10)
    out["local_observations_per_sec"] = Bo / (ms * 1e-3)
    out["map_point_tests_per_sec_algorithmic"] = Bo * 729.0 * M / (ms * 1e-3)
    try:
        from crl_kino_b200.estimator import load_estimator
        w = np.load(os.path.join(ROOT, "data", "estimator_weights.npz"))
        est, cls = load_estimator({k[4:]: w[k] for k in w.files if k.startswith("est.")},
                                  {k[4:]: w[k] for k in w.files if k.startswith("cls.")})
        _, o32, _ = ops.local_obs(env.handle, states[:Bo], goals[:Bo], want64=False, want32=True)
        ms = timed(lambda: ops.estimator_forward(est.pair.handle, o32), 10)
        out["estimator_pairs_per_sec"] = Bo / (ms * 1e-3)
        out["estimator_tflops_algorithmic"] = 2 * 1.158038e6 * Bo / (ms * 1e-3) / 1e12     # 2 nets x 579,019 MAC x 2
    except Exception as e:                                                     # weights file not present
        out["estimator"] = f"skipped: {e}"
    Bg = 1 << 15
    gs = states[:Bg].clone()
    acts = ops.as_dev(np.stack([rng.uniform(-0.1, 1.0, Bg), rng.uniform(-np.pi, np.pi, Bg)], 1))
    rp = np.array([50.0, -1.5, -0.5, -300.0, 1.0, 1.0, -.5, -2.5])
    ms = timed(lambda: ops.gym_step(env.handle, gs, goals[:Bg], acts, rp, 0.2, want64=False, want32=True), 5)
    out["gym_env_steps_per_sec"] = Bg / (ms * 1e-3)
    Bd = 4096
    free = states[:Bd].clone()
    free[:, 3] = 0.4
    free[:, 4] = 0.0
    r = ops.dwa_control(env.handle, dwa.params(), free, goals[:Bd])
    ms = timed(lambda: ops.dwa_control(env.handle, dwa.params(), free, goals[:Bd]), 5)
    n_roll = float(r["n_rollouts"].double().sum().item())
    out["dwa_controls_per_sec"] = Bd / (ms * 1e-3)
    out["rollouts_per_sec"] = n_roll / (ms * 1e-3)
    out["rollout_tflops_algorithmic"] = M_FLOP_ROLLOUT * n_roll / (ms * 1e-3) / 1e12
    return out


# ------------------------------------------------------------------ cfg5 leg

def cfg5_leg(args, env, rank, local_rank, world, barrier):
    """BASELINE config 5: --cfg5-total independent queries sharded over the ranks (contiguous blocks,
    planner/sharding.py), --cfg5-nodes-node trees, batched policy steering.  One step = nearest node of
    one sample per query (k_nearest streams the float32 coordinates: 8 B per node) + one RL extension
    (<= 50 policy steps, rrt_rl.py:22-63) from that node toward the query's goal.  No collective on
    the data path; after the last step the per-query results are gathered on rank 0 over NCCL
    (sharding.gather_results), timed separately."""
    import torch
    import torch.distributed as dist
    from crl_kino_b200 import ops
    from crl_kino_b200.env import DifferentialDriveGym
    from crl_kino_b200.planner.batched import BatchedExtender, Forest
    from crl_kino_b200.planner.sharding import gather_results, shard_range
    from crl_kino_b200.policy.rl_policy import load_policy
    from crl_kino_b200.workloads import free_states

    N = args.cfg5_nodes
    total_q = args.cfg5_total
    lo, hi = shard_range(total_q, rank, world)
    Q = hi - lo
    torch.cuda.empty_cache()
    free_b, total_b = torch.cuda.mem_get_info()
    per_query = N * 48 + 50 * 5 * 8 * 3 + 4096                 # forest node = 40 B state + 8 B float32 xy; path + staging
    reserve = 6 << 30                                            # RL workspace (8192-query passes), NCCL, allocator slack
    fit = int((free_b - reserve) / per_query)
    capped = Q > fit
    Q = max(min(Q, fit), 1)
    pol = load_policy(DifferentialDriveGym(), [1024, 512, 512, 512], None, seed=0)
    rng = np.random.default_rng(3 + rank)

    def clearance(c):
        return env.valid_state_batch(c, want_clearance=True)[1].cpu().numpy()
    pool = free_states(clearance, rng, 1 << 20, 0.0, batch=1 << 19)
    pool[:, 3:] = 0.0
    dpool = ops.as_dev(pool)
    forest = Forest(Q, N, with_xy64=False, with_parent=False)
    g = torch.Generator(device="cuda").manual_seed(11 + rank)
    for q0 in range(0, Q, 64):
        q1 = min(Q, q0 + 64)
        idx = torch.randint(0, dpool.shape[0], (q1 - q0, N), generator=g, device="cuda")
        st = dpool[idx]
        forest.states[q0:q1] = st
        forest.x32[q0:q1] = st[:, :, 0].float(); forest.y32[q0:q1] = st[:, :, 1].float()
        del st, idx
    forest.n_nodes[:] = N
    K, W = args.cfg5_steps, args.cfg5_warmup
    samples = [ops.as_dev(free_states(clearance, rng, Q, 0.0, batch=1 << 17)) for _ in range(K + W)]
    goals = ops.as_dev(free_states(clearance, rng, Q, 0.5, batch=1 << 17))
    ext = BatchedExtender(env, forest, policy=pol, eps=0.0, rl_chunk=8192, pipe_depth=2)
    ext.extend_targets = goals
    sampler = ClockSampler(local_rank)
    sampler.wait_first()
    t_load0 = time.perf_counter()
    for w in range(W):
        ext.step(samples[w])
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    barrier()
    policy_evals = 0
    for k in range(K):
        out = ext.step(samples[W + k], events=evs[k])
        policy_evals += out["n_steps"].sum(dtype=torch.int64)
    barrier()
    clock_windows = [(t_load0, time.perf_counter())]
    t_near = sum(e[0].elapsed_time(e[1]) for e in evs)
    t_ext = sum(e[2].elapsed_time(e[3]) for e in evs)
    step_ms = [e[0].elapsed_time(e[3]) for e in evs]
    t_total = sum(step_ms)
    policy_evals = float(policy_evals.item())
    mean_steps = float(out["n_steps"].double().mean().item())
    st = out["status"].cpu().numpy()
    # ---- the final result gather (the only collective): per-query outcome rows onto rank 0
    last = torch.clamp(out["n_steps"].long() - 1, min=0)
    end_state = out["path"][torch.arange(Q, device="cuda"), last]
    local = {"status": out["status"], "n_steps": out["n_steps"], "n_nodes": out["n_nodes"],
             "nearest_idx": out["nearest_idx"], "end_state": end_state}
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gather_ms, gathered_rows, gather_ok, gather_bytes = 0.0, Q, True, 0
    if world > 1 and not capped:
        gather_results(local, total_q)                       # warm-up (NCCL channel setup)
        barrier()
        g0.record()
        res = gather_results(local, total_q)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)
        if rank == 0:
            gathered_rows = int(res["status"].shape[0])
            gather_ok = bool(gathered_rows == total_q and torch.equal(res["n_steps"][lo:hi], out["n_steps"]) and
                             torch.equal(res["end_state"][lo:hi], end_state))
            gather_bytes = int(sum(v.numel() * v.element_size() for v in res.values()))
    # ---- end to end: samples from pinned host memory, results back to the host, every step
    hs = [s.cpu().numpy() for s in samples]
    t_load0 = time.perf_counter()
    ext.collect(ext.submit_host(hs[0]))
    ext.collect(ext.submit_host(hs[1 % len(hs)]))
    barrier()
    t0 = time.perf_counter()
    pending = None
    n_back = 0
    for k in range(K):                                   # two steps in flight: a step's D2H under the next step's kernels
        t = ext.submit_host(hs[W + k])
        if pending is not None:
            n_back += int(ext.collect(pending)["n_steps"].shape[0])
        pending = t
    n_back += int(ext.collect(pending)["n_steps"].shape[0])
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    assert n_back == K * Q
    clock_windows.append((t_load0, time.perf_counter()))
    clocks = sampler.stop(clock_windows)
    clocks["window"] = "warm-up + timed steps and the end-to-end loop of this leg (same step, same load)"
    red = torch.tensor([t_total, t_e2e, t_near, t_ext, gather_ms], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(Q * K), policy_evals, float(Q)], dtype=torch.float64, device="cuda")
    per_rank = [t_total / K]
    if world > 1:
        allr = [torch.zeros_like(red) for _ in range(world)]
        dist.all_gather(allr, red)
        per_rank = [float(a[0]) / K for a in allr]
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    t_total, t_e2e, t_near_m, t_ext_m, gather_ms = [float(x) for x in red.cpu()]
    tot, evals, tot_q = [float(x) for x in cnt.cpu()]
    h2d, d2h = ext.h2d_bytes(), ext.d2h_bytes()
    launches, bpn = ext.launches_per_step, forest.bytes_per_node()
    del ext, forest, samples, goals, dpool, out, local, end_state
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peaks = load_peaks()
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    bf16 = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
    near_gbs = 8.0 * Q * N * K / (t_near * 1e-3) / 1e9
    rl_tflops = POLICY_FLOP * (evals / world) / (t_ext_m * 1e-3) / 1e12
    ncu = load_ncu("r2_mlp_chain_ncu.json")
    ncu_adv = load_ncu("r2_extend_rl_ncu.json")
    return {
        "metric": "rrt_extensions_per_sec", "value": tot / (t_total * 1e-3), "unit": "extensions/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_total / K, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32 policy (bf16 split operands on tcgen05) / f64 dynamics",
        "data": "synthetic",
        "config": {"workload": "cfg5: 65,536 independent queries sharded over the GPUs, 100,000-node trees, batched policy steering",
                   "queries_total": int(tot_q), "queries_per_gpu": Q, "capped_by_memory": bool(capped), "tree_nodes": N,
                   "bytes_per_node": bpn, "obstacles": len(env.rects),
                   "policy": "seeded random-init actor 733-1024-512-512-512-2, eps 0 (the trained policy.pth is missing upstream)",
                   "l2": "per-step tree scan (8 B x nodes x queries per GPU) exceeds L2"},
        "workload_stats": {"mean_policy_steps_per_extension": mean_steps,
                           "status_last_step_rank0": {"timeout": int((st == 0).sum()), "collision": int((st == 1).sum()),
                                                      "reached": int((st == 2).sum())}},
        "e2e": {"value": tot / t_e2e, "unit": "extensions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "BatchedExtender.submit_host / collect (two steps in flight: a step's D2H under the next step's kernels)"},
        "per_rank_ms": per_rank, "step_ms": {"min": min(step_ms), "median": float(np.median(step_ms)), "max": max(step_ms)},
        "kernel_ms_per_step": {"nearest": t_near_m / K, "extend_rl": t_ext_m / K},
        "gather_ms": gather_ms, "gather": {"collective": "torch.distributed.gather over NCCL (sharding.gather_results)",
                                           "rows_on_rank0": gathered_rows, "bytes_on_rank0": gather_bytes,
                                           "rank0_shard_intact": gather_ok,
                                           "skipped": bool(world == 1 or capped)},
        "gpu_launches": K * launches,
        "roofline": {"kernel": "extend_rl = per policy step k_mlp_chain (the actor's hidden layers, ~58 % of the pass) + "
                               "k_rl_advance (head / motion / validity / bookkeeping / next observation)",
                     "bound": "tensor", "achieved": rl_tflops, "peak": bf16, "unit": "TFLOP/s", "frac": rl_tflops / bf16,
                     "traffic": ncu.get("dram_bytes") if ncu else None, "ncu": ncu, "ncu_step_kernel": ncu_adv,
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else "fallback",
                     "policy_evaluations_per_step_per_gpu": evals / world / K,
                     "note": "algorithmic float32 FLOP of the actor (3,600,384 per policy evaluation) over the executed "
                             "policy steps; the operand split multiplies the tensor-pipe work (DESIGN.md section 3)"},
        "rooflines_other_kernels": [{"kernel": "k_nearest", "bound": "hbm", "achieved": near_gbs, "peak": hbm,
                                     "unit": "GB/s", "frac": near_gbs / hbm, "traffic": None}],
        "clocks": clocks}


# ------------------------------------------------------------------ our arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_info()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from crl_kino_b200.env import DifferentialDriveEnv
    from crl_kino_b200.workloads import dense_map
    rects = dense_map(n_cells=args.cells)
    env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi)
    env.set_obs(rects)
    line = None
    if args.workload in ("auto", "cfg4"):
        line = cfg4_leg(args, env, rects, rank, local_rank, world, barrier)
    if args.workload == "cfg5" or (args.workload == "auto" and world > 1):
        c5 = cfg5_leg(args, env, rank, local_rank, world, barrier)
        if rank == 0:
            if line is None:
                line = c5
            else:
                line["cfg5"] = c5
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and line is not None:
        print(json.dumps(line))


def cfg4_leg(args, env, rects, rank, local_rank, world, barrier):
    import torch
    import torch.distributed as dist
    from crl_kino_b200 import ops
    from crl_kino_b200.planner.batched import BatchedExtender, Forest
    from crl_kino_b200.policy.dwa import DWA
    from crl_kino_b200.workloads import DENSE_DWA, cfg4_workload
    args.queries = args.queries or 1024
    dwa = DWA(env)
    dwa.set_dwa(**DENSE_DWA)
    if args.use_clearance:
        dwa.set_dwa(use_clearance=True)
        args.no_cpu = True          # the CPU restatement needs ~80 s per control step in this mode
        args.no_extra = True
    K, W, Q = args.steps, args.warmup, args.queries
    wl = cfg4_workload(Q=Q, tree_nodes=args.tree_nodes, n_batches=K + W, seed=WORKLOAD_SEED + rank, cells=args.cells)
    batches = wl["batches"]
    forest = Forest(Q, args.tree_nodes)
    forest.fill(wl["trees"])
    ext = BatchedExtender(env, forest, dwa=dwa, pipe_depth=args.pipe_depth)
    dev_batches = ops.as_dev(batches)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    # ---- device-resident arm ------------------------------------------------
    for w in range(W):
        ext.step(dev_batches[w])
    barrier()
    ext.stats.zero_()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    sampler = ClockSampler(local_rank)
    sampler.wait_first()
    barrier()
    torch.cuda.profiler.start()          # ncu --profile-from-start off captures only the timed region
    t_load0 = time.perf_counter()
    t_wall = time.time()
    for k in range(K):
        flush.fill_(k & 0xff)                      # L2 flush between timed iterations (untimed)
        ext.step(dev_batches[W + k], events=evs[k])
    barrier()
    t_wall = time.time() - t_wall
    torch.cuda.profiler.stop()
    t_near = sum(e[0].elapsed_time(e[1]) for e in evs)
    t_gath = sum(e[1].elapsed_time(e[2]) for e in evs)
    t_ext = sum(e[2].elapsed_time(e[3]) for e in evs)
    step_ms = [e[0].elapsed_time(e[3]) for e in evs]
    t_total_ms = t_near + t_gath + t_ext
    stats = ext.stats.cpu().numpy().astype(np.float64)
    n_ext, n_ctrl, n_roll = stats[0], stats[1], stats[2]
    status = ext.out["status"].cpu().numpy()

    # ---- sustained: the same step loop (all batches, cyclic, flush between steps) for >= --sustain-s
    sustained = None
    if args.sustain_s > 0:
        chunk = 256
        sev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(chunk)]
        ms_sum, n_done, t0 = 0.0, 0, time.perf_counter()
        while time.perf_counter() - t0 < args.sustain_s:
            for i in range(chunk):
                flush.fill_(i & 0xff)
                sev[i][0].record()
                ext.step(dev_batches[(n_done + i) % (K + W)])
                sev[i][1].record()
            torch.cuda.synchronize()
            ms_sum += sum(a.elapsed_time(b) for a, b in sev)
            n_done += chunk
        sustained = {"steps": n_done, "ms_per_step": ms_sum / n_done, "wall_s": time.perf_counter() - t0,
                     "value_rank0": Q * n_done / (ms_sum * 1e-3),
                     "note": "same loop as the timed region over all K + W batches, cyclic, L2 flushed between steps; "
                             "per-step CUDA-event time summed"}

    # ---- the same K timed batches with pipe_depth steps in flight, inputs and results resident in HBM:
    # one launch of 1024 extensions lasts as long as its longest extension (tail bound), the next steps'
    # extensions take the SMs the tail leaves idle.  The L2 flush runs on each step's stream before the step.
    pipelined = None
    if args.sustain_s > 0:
        for w in range(W):
            ext.wait(ext.submit(dev_batches[w], flush=flush))
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        tickets = [ext.submit(dev_batches[W + k], flush=flush) for k in range(K)]
        for t in tickets[-ext.pipe_depth:]:
            ext.wait(t)
        p1.record()
        torch.cuda.synchronize()
        pipelined = {"ms_per_step": p0.elapsed_time(p1) / K, "value_rank0": Q * K / (p0.elapsed_time(p1) * 1e-3),
                     "steps_in_flight": ext.pipe_depth,
                     "note": "device-resident BatchedExtender.submit / wait over the K timed batches, L2 flush (256 MiB "
                             "write) on every step's stream inside the timed region; explains e2e > value: `value` times "
                             "one launch at a time"}
    clocks = sampler.stop([(t_load0, time.perf_counter())])
    clocks["window"] = "timed region + the sustained loop that follows it (same step loop)"

    # ---- end-to-end arm: host buffers through the public API -------------------
    # submit_host / collect: every step stages its samples in pinned memory, copies them to the
    # device, runs nearest + gather + extension and copies path / status / nodes back; up to three
    # steps are in flight, each on its own stream, so that the D2H of a step overlaps the kernels of the
    # next and the next step's extensions fill the SMs that the tail of a step leaves idle.
    from collections import deque
    for w in range(W):
        ext.collect(ext.submit_host(batches[w]))
    barrier()
    t0 = time.perf_counter()
    pending = deque()
    n_checked = 0
    for k in range(K):
        pending.append(ext.submit_host(batches[W + k]))
        if len(pending) == ext.pipe_depth:             # oldest step's results (pinned host memory)
            res = ext.collect(pending.popleft())
            n_checked += int(res["n_steps"].shape[0])
    while pending:
        res = ext.collect(pending.popleft())
        n_checked += int(res["n_steps"].shape[0])
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    assert n_checked == K * Q
    t0 = time.perf_counter()
    for k in range(K):
        ext.step_host(batches[W + k])               # the synchronous call (one step at a time), for comparison
    t_e2e_sync = time.perf_counter() - t0

    # ---- the results of timed batch 0 with the chosen rollout of every control step (parity leg)
    gpu0 = None
    if rank == 0 and not args.no_cpu:
        f = forest
        idx, dist0, _ = f.nearest(dev_batches[W], ext.coord_bound)
        froms = ops.gather_rows(f.states, idx)
        o = ops.extend_dwa(env.handle, dwa.params(), ext.xp, froms, dev_batches[W], want_ctrl=True)
        gpu0 = {k: v.cpu().numpy() for k, v in o.items() if v is not None}
        gpu0["nearest_idx"] = idx.cpu().numpy()

    plan = None
    if not args.no_extra:
        plan = planning_leg(args, env, dwa, rank)

    # ---- reduce over ranks (max time, summed work) ----------------------------------
    mine = torch.tensor([t_total_ms, t_e2e, t_ext, plan["seconds"] if plan else 0.0,
                         plan["e2e_seconds"] if plan else 0.0, sustained["ms_per_step"] if sustained else 0.0,
                         clocks["sm_mhz"] or 0.0, t_near, max(step_ms)], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([n_ext, n_ctrl, n_roll, plan["queries"] if plan else 0.0, plan["extensions"] if plan else 0.0],
                       dtype=torch.float64, device="cuda")
    red = mine.clone()
    per_rank = [mine.cpu().tolist()]
    if world > 1:
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [a.cpu().tolist() for a in allr]
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    t_total_max, t_e2e_max, t_ext_max, t_plan, t_plan_e2e, sus_max = [float(x) for x in red.cpu()[:6]]
    tot_ext, tot_ctrl, tot_roll, tot_q, tot_pext = [float(x) for x in cnt.cpu()]
    if rank != 0:
        del ext, forest, dev_batches, flush
        torch.cuda.empty_cache()
        return None

    peaks = load_peaks()
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    fp64_peak = sms * FP64_LANES_PER_SM * 2 * sm_max * 1e6 / 1e12
    M = len(rects)
    flop = M_FLOP_ROLLOUT * n_roll + M_FLOP_PAIR * M * n_ctrl          # this rank's extend launches, SURVEY 8d count
    achieved = flop / (t_ext * 1e-3) / 1e12
    executed = None
    if len(stats) >= 8 and stats[5] > 0:
        xf = (X_FLOP["rollout"] * n_roll + X_FLOP["sincos_task"] * stats[5] + X_FLOP["velocity_chain"] * stats[6] +
              X_FLOP["pair_test"] * stats[3] + X_FLOP["step_fixed"] * n_ctrl)
        executed = {"fp64_flop_per_launch": xf / K, "tflops": xf / (t_ext * 1e-3) / 1e12,
                    "counters_per_launch": {"control_steps": n_ctrl / K, "rollouts": n_roll / K,
                                            "exact_pair_tests": stats[3] / K, "rectangles_screened_fp32": stats[4] / K,
                                            "sincos_chain_tasks": stats[5] / K, "velocity_chains": stats[6] / K},
                    "per_unit_fp64_flop": X_FLOP,
                    "source": "counters maintained by k_extend_dwa (crlk_extend_dwa `stats`) x per-unit FP64 operation "
                              "counts (DESIGN.md section 4); frac_executed counts an FMA as 2 against the FMA peak, the "
                              "pipe's busy share is in roofline.ncu (fp64_pipe_pct_while_active)"}
    # the ncu summary belongs to the default kernel; the clearance variant (k_extend_dwa<.., true>) has its own
    # file when one was captured, and its per-rollout clearance searches are not in the work counters
    ncu = load_ncu("r2_extend_clr_ncu.json" if args.use_clearance else "r2_extend_ncu.json")
    if args.use_clearance:
        executed = None
    value = tot_ext / (t_total_max * 1e-3)
    line = {
        "metric": "rrt_extensions_per_sec", "value": value, "unit": "extensions/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_total_max / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg4_config(args, M),
        "workload_stats": {"fingerprint_rank0": workload_fingerprint(wl, W),
                           "rollouts_per_control_realised": tot_roll / max(tot_ctrl, 1),
                           "control_steps_per_extension": tot_ctrl / max(tot_ext, 1),
                           "control_steps_per_sec": tot_ctrl / (t_total_max * 1e-3),
                           "rollouts_per_sec": tot_roll / (t_total_max * 1e-3),
                           "status_last_step_rank0": {"timeout": int((status == 0).sum()), "collision": int((status == 1).sum()),
                                                      "reached": int((status == 2).sum())}},
        "e2e": {"value": tot_ext / t_e2e_max, "unit": "extensions/s",
                "h2d_bytes_per_step": ext.h2d_bytes(), "d2h_bytes_per_step": ext.d2h_bytes(),
                "api": f"BatchedExtender.submit_host / collect ({ext.pipe_depth} independent steps in flight, one stream each)",
                "one_step_at_a_time_rank0": Q * K / t_e2e_sync},
        "gpu_launches": K * ext.launches_per_step,
        "kernel_ms_per_step": {"nearest": t_near / K, "gather": t_gath / K, "extend_dwa": t_ext / K},
        "step_ms": {"min": min(step_ms), "median": float(np.median(step_ms)), "max": max(step_ms)},
        "per_rank": [{"rank": r, "ms_per_step": p[0] / K, "extend_ms_per_step": p[2] / K, "nearest_ms_per_step": p[7] / K,
                      "max_step_ms": p[8], "e2e_s": p[1], "sustained_ms_per_step": p[5], "sm_mhz": p[6]}
                     for r, p in enumerate(per_rank)],
        "pipelined_device_resident": pipelined,
        "sustained": None if sustained is None else dict(sustained, value=world * Q / (sus_max * 1e-3),
                                                         ms_per_step_max_over_ranks=sus_max),
        "roofline": {"kernel": "k_extend_dwa", "bound": "fp64", "achieved": achieved, "peak": fp64_peak,
                     "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                     "frac_kind": "algorithmic: SURVEY 8d counts 406 FLOP per rollout + 860 per (pose, obstacle) pair of a "
                                  "brute-force scan; the kernel never executes most of it (shared sincos chains, grid + "
                                  "FP32 screen before the exact FP64 pair test), so this is NOT a hardware fraction",
                     "frac_executed": None if executed is None else executed["tflops"] / fp64_peak,
                     "executed": executed,
                     "traffic": ncu.get("dram_bytes") if ncu else None, "ncu": ncu,
                     "peak_source": f"derived: {sms} SM x {FP64_LANES_PER_SM} FP64 FMA lanes x 2 x "
                                    f"{sm_max:.0f} MHz (MEASURED_PEAKS.json holds only HBM and bf16 peaks)",
                     "algorithmic_flop_per_launch": flop / K},
        "clocks": clocks, "wall_s_timed_region": t_wall,
    }
    if plan is not None:
        line["planning"] = {"queries_per_sec": tot_q / t_plan, "unit": "queries/s", "queries": int(tot_q),
                            "max_iter": plan["max_iter"], "ms": t_plan * 1e3,
                            "e2e": {"queries_per_sec": tot_q / t_plan_e2e, "h2d_bytes": plan["h2d_bytes"],
                                    "d2h_bytes": plan["d2h_bytes"]},
                            "sampling": "on the device (counter-based generator = RRT.sample, rrt.py:146-154)",
                            "reached_frac_rank0": plan["reached_frac"], "mean_nodes_rank0": plan["mean_nodes"],
                            "mean_samples_consumed_rank0": plan["mean_samples"],
                            "max_samples_consumed_rank0": plan["max_samples_consumed"],
                            "stopped_on_sample_limit_rank0": plan["stopped_on_sample_limit"],
                            "extensions_per_sec": tot_pext / t_plan, "gpu_launches": 1,
                            "from_host_streams_rank0": {"queries_per_sec": plan["queries"] / plan["streams_e2e_seconds"],
                                                        "h2d_bytes": plan["streams_h2d_bytes"],
                                                        "streams_exhausted": plan["streams_exhausted"],
                                                        "note": "pre-drawn sample streams from pinned host memory (the "
                                                                "parity tests' form), wall clock incl. H2D and D2H"},
                            "note": "Q full RRT.planning runs in one launch of k_plan_dwa (FusedRRT: one CTA per "
                                    "query from first sample to last node, persistent CTAs pull queries); CUDA-event "
                                    "time, best of 2 after warm-up; e2e adds the H2D of the queries and the D2H of the results"}
        if "lockstep_seconds" in plan:
            line["planning"]["lockstep"] = {"queries_per_sec_rank0": plan["queries"] / plan["lockstep_seconds"],
                                            "rounds": plan["lockstep_rounds"],
                                            "identical_results": plan["lockstep_identical"],
                                            "note": "BatchedRRT: five launches per round, wall clock"}
        line["rooflines_other_kernels"] = [nearest_roofline(peaks), mlp_roofline(peaks)]
        line["sub_rates"] = sub_rates(env, dwa, peaks)
        line["planning_rl"] = planning_rl_leg(args, env, rects, args.no_cpu)
    h2d, d2h = ext.h2d_bytes(), ext.d2h_bytes()
    del ext, forest, dev_batches, flush
    torch.cuda.empty_cache()
    if not args.no_cpu:
        from oracle import oracle as orc
        cores = orc.cpu_count()
        n = args.cpu_sample or Q
        v, dt, ref = cpu_sample_run(rects, wl["trees"], batches[W], n, cores, want_ctrl=True)
        line["cpu_baseline"] = {"value": v, "unit": "extensions/s", "cores": cores, "kind": "port",
                                "control_steps_per_extension": float(ref["n_steps"].sum()) / n,
                                "sample": f"first {n} queries of timed batch 0 (nearest + full DWA extension, "
                                          f"{int(ref['n_steps'].sum())} control steps) through oracle/crlk_oracle.c on "
                                          f"{cores} threads in {dt:.1f} s"}
        line["parity"] = parity_block(gpu0, ref, n)
        if plan is not None:
            ncpu = len(plan["inputs"][0])
            cp = cpu_planning(rects, plan["inputs"], plan["max_iter"], ncpu, cores)
            g = plan["gpu_first"]
            cp["same_outcome_as_gpu"] = int(sum(1 for q in range(ncpu) if cp["_nodes"][q] == int(g["nodes"][q]) and
                                                cp["_reached"][q] == bool(g["reached"][q])))
            del cp["_nodes"], cp["_reached"]
            line["planning"]["cpu_baseline"] = cp
        if not args.no_pyref:
            line["cpu_baseline"]["python_reference"] = python_reference_cfg1()
    return line


# ------------------------------------------------------------------ reference arm

def run_reference(args):
    """--impl reference: the CPU restatement of the reference path (oracle/crlk_oracle.c; the Python
    reference needs ~0.5 s per control step on this map) on all host threads, on the SAME workload as
    our arm (rank 0's): every step runs a bounded prefix of that step's batch."""
    rank, _, world = dist_info()
    if rank != 0:
        return
    args.queries = args.queries or 1024
    from crl_kino_b200.workloads import cfg4_workload, dense_map
    from oracle import oracle as orc
    rects = dense_map(n_cells=args.cells)
    cores = orc.cpu_count()
    K, W = args.steps, args.warmup
    n = args.cpu_sample or min(args.queries, max(16 * cores, 128))
    full = cfg4_workload(Q=max(n, 64), tree_nodes=args.tree_nodes, n_batches=K + W, seed=WORKLOAD_SEED, cells=args.cells)
    trees, batches = full["trees"], full["batches"]
    for w in range(min(W, 1)):
        cpu_sample_run(rects, trees, batches[w], n, cores)
    t0 = time.perf_counter()
    tot, ctl = 0, 0
    for k in range(K):
        _, _, r = cpu_sample_run(rects, trees, batches[W + k], n, cores)
        tot += n
        ctl += int(r["n_steps"].sum())
    dt = time.perf_counter() - t0
    v = tot / dt
    line = {
        "impl": "reference", "metric": "rrt_extensions_per_sec", "value": v, "unit": "extensions/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg4_config(args, len(rects)),
        "workload_stats": {"fingerprint_rank0": workload_fingerprint(full, W),
                           "control_steps_per_extension": ctl / max(tot, 1), "control_steps_per_sec": ctl / dt},
        "cpu_baseline": {"value": v, "unit": "extensions/s", "cores": cores, "kind": "port",
                         "sample": f"first {n} of the {args.queries} queries of every batch (same batches as the GPU arm's "
                                   f"rank 0), nearest + full DWA extension through oracle/crlk_oracle.c on {cores} threads"},
        "e2e": {"value": v, "unit": "extensions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.no_pyref:
        line["cpu_baseline"]["python_reference"] = python_reference_cfg1()
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.pyref:
        pyref_child(*a.pyref)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
