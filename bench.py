#!/usr/bin/env python
"""bench.py -- RRT extensions/s of the batched extend path on B200.

One "step" = one pass of the hot path over one batch: for each of Q queries,
nearest node of its sample over the query's tree, then one whole extension
(<= 50 closed-loop DWA control steps with motion + collision checks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); queries are independent,
so ranks share nothing but the final reduction of the counters ("weak" scaling:
every rank runs the full per-GPU workload).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_FLOP_ROLLOUT = 406.0      # FP ops per DWA rollout
M_FLOP_PAIR = 860.0         # FP ops per (pose, obstacle) pair, brute force
FP64_LANES_PER_SM = 64      # B200: 64 FP64 FMA lanes per SM
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures
# summarised in profiles/r1d_* / r1h_*_ncu_summary.txt (default workload sizes only)
NCU_TRAFFIC_EXTEND_DWA = 244.224e3 + 0.0               # k_extend_dwa, 1024 queries, cfg4 map (profiles/r1t_extend_ncu_summary.txt)
NCU_TRAFFIC_NEAREST = 820.148992e6 + 4.725248e6       # k_nearest, 1024 trees x 100,000 nodes


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg5"])
    ap.add_argument("--queries", type=int, default=0, help="queries per GPU (0 = 1024 for cfg4, 65536 / gpus for cfg5)")
    ap.add_argument("--tree-nodes", type=int, default=1024)
    ap.add_argument("--cells", type=int, default=10000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="extensions in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--use-clearance", action="store_true",
                    help="enable the per-rollout clearance term (dwa.py:64-69); GPU legs only")
    ap.add_argument("--no-extra", action="store_true", help="skip the planning / nearest / MLP side measurements")
    ap.add_argument("--plan-iters", type=int, default=100, help="max_iter of the planning-queries/s leg")
    ap.add_argument("--plan-queries", type=int, default=0, help="queries per GPU of the planning leg (0 = --queries)")
    ap.add_argument("--plan-lockstep", action="store_true", help="also time the lock-step planner (BatchedRRT)")
    return ap.parse_args()


def dist_info():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs."""
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
            self.rows.append(line.strip())

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
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


def build_workload(args, env, rank):
    """cfg4: dense 10,004-obstacle map, Q queries with pre-filled trees of free
    nodes, and (steps + warmup) batches of accepted samples (valid state,
    nearest-node distance in [1, 20), rrt.py:66-69)."""
    import torch
    from crl_kino_b200 import ops
    from crl_kino_b200.planner.batched import Forest
    from crl_kino_b200.workloads import free_states, random_states
    rng = np.random.default_rng(100 + rank)
    Q, N = args.queries, args.tree_nodes

    def clearance(c):
        return env.valid_state_batch(c, want_clearance=True)[1].cpu().numpy()
    nodes = free_states(clearance, rng, Q * N, 0.5, batch=1 << 18)
    nodes[:, 3:] = 0.0
    forest = Forest(Q, N)
    forest.fill(nodes.reshape(Q, N, 5))
    nb = args.steps + args.warmup
    batches = np.zeros((nb, Q, 5))
    need = np.ones((nb, Q), dtype=bool)
    while need.any():
        cand = random_states(rng, nb * Q, moving=True).reshape(nb, Q, 5)
        v = env.valid_state_batch(cand.reshape(-1, 5))[0].cpu().numpy().astype(bool).reshape(nb, Q)
        for b in range(nb):
            _, dist, _ = ops.nearest(forest.x, forest.y, forest.n_nodes, cand[b], x32=forest.x32,
                                     y32=forest.y32, coord_bound=env.coord_bound())
            d = dist.cpu().numpy()
            ok = need[b] & v[b] & (d >= 1.0) & (d < 20.0)
            batches[b, ok] = cand[b, ok]
            need[b, ok] = False
    return forest, batches


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


def planning_leg(args, env, dwa, rank):
    """Planning queries/s: Q full RRT.planning runs (max_iter = --plan-iters).
    fused    = FusedRRT (crlk_plan_dwa, one launch, one CTA per query): device time by CUDA
               events, and end to end with the streams coming from pinned host memory and
               the per-query results read back;
    lockstep = BatchedRRT (five launches per round, all queries in lock step), wall clock."""
    import torch
    from crl_kino_b200.planner.batched import BatchedRRT, FusedRRT
    from crl_kino_b200.workloads import query_pairs
    Q = args.plan_queries or args.queries

    def clearance(c):
        return env.valid_state_batch(c, want_clearance=True)[1].cpu().numpy()
    starts, goals = query_pairs(clearance, Q, seed=7 + rank)
    sb = env.get_bounds()["state_bounds"]
    S = args.plan_iters * 4 + 32
    rng = np.random.default_rng(300 + rank)
    streams = rng.uniform(sb[:, 0], sb[:, 1], (Q, S, 5))
    pick = rng.integers(0, 101, (Q, S)) <= 5                      # goal bias of rrt.py:147
    streams[pick] = np.repeat(goals[:, None, :], S, 1)[pick]
    h_streams = torch.as_tensor(streams).pin_memory()
    dstreams = h_streams.cuda()
    out = {"queries": Q, "max_iter": args.plan_iters, "inputs": (starts, goals, streams)}

    fp = FusedRRT(env, starts, goals, max_iter=args.plan_iters, dwa=dwa, keep_paths=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for rep in range(3):                       # first pass = warm-up, best of the next two
        fp.reset()
        torch.cuda.synchronize()
        if rep == 2:
            torch.cuda.profiler.start()        # second ncu range (the first is the extend step loop)
        e0.record()
        fp.plan_async(dstreams)
        e1.record()
        torch.cuda.synchronize()
        if rep == 2:
            torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1)
        if rep > 0:
            best = ms if best is None else min(best, ms)
    st = fp.stats.cpu().numpy()
    reached = fp.reached.cpu().numpy().astype(bool)
    exhausted = int(((fp.cursor.cpu().numpy() >= S) & ~reached & (fp.iters.cpu().numpy() < args.plan_iters)).sum())
    out.update(seconds=best * 1e-3, reached_frac=float(reached.mean()), extensions=int(st[0]),
               control_steps=int(st[1]), mean_nodes=float(fp.forest.n_nodes.double().mean().item()),
               mean_samples=float(fp.cursor.double().mean().item()), streams_exhausted=exhausted)
    # end to end: pinned host streams in, per-query results out, inside the timed region
    d_in = torch.empty_like(dstreams)
    h_res = [torch.empty(Q, dtype=torch.uint8).pin_memory(), torch.empty(Q, dtype=torch.int32).pin_memory(),
             torch.empty(Q, dtype=torch.int32).pin_memory()]
    fp.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d_in.copy_(h_streams, non_blocking=True)
    fp.plan_async(d_in)
    for h, t in zip(h_res, (fp.reached, fp.iters, fp.forest.n_nodes)):
        h.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    out["e2e_seconds"] = time.perf_counter() - t0
    out["h2d_bytes"] = h_streams.numel() * 8
    out["d2h_bytes"] = sum(h.numel() * h.element_size() for h in h_res)
    assert np.array_equal(h_res[0].numpy().astype(bool), reached)
    nodes_from_streams = fp.forest.n_nodes.clone()
    # the same planner drawing its own samples on the device (no sample streams at all)
    fp.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fp.plan_async(None, seed=300 + rank, max_samples=S)
    for h, t in zip(h_res, (fp.reached, fp.iters, fp.forest.n_nodes)):
        h.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    out["device_sampling_seconds"] = time.perf_counter() - t0
    out["device_sampling_reached_frac"] = float(h_res[0].numpy().astype(bool).mean())

    if args.plan_lockstep:
        p = BatchedRRT(env, starts, goals, max_iter=args.plan_iters, dwa=dwa, keep_paths=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r2 = p.plan(dstreams, check_every=16)
        out["lockstep_seconds"] = time.perf_counter() - t0
        out["lockstep_rounds"] = p.samples_used
        out["lockstep_identical"] = bool(np.array_equal(r2, reached) and
                                         torch.equal(p.forest.n_nodes, nodes_from_streams))
    return out


def _cpu_plan_one(a):
    rects, start, goal, stream, iters = a
    from crl_kino_b200.workloads import DENSE_DWA
    from oracle import oracle as orc
    o = orc.Env()
    o.set_obs(rects)
    p = orc.PlannerPort(o, max_iter=iters, dwa_kwargs=DENSE_DWA)
    p.set_start_and_goal(start, goal)
    p.planning(iter(stream))
    return p.n_extend_calls, bool(p.reach_exactly)


def cpu_planning(rects, inputs, iters, n, cores):
    import multiprocessing as mp
    starts, goals, streams = inputs
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_plan_one, [(rects, starts[q], goals[q], streams[q], iters) for q in range(n)])
    dt = time.perf_counter() - t0
    return {"queries_per_sec": n / dt, "queries": n, "seconds": dt, "cores": cores, "kind": "port",
            "extensions": int(sum(r[0] for r in res))}


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
    return {"kernel": "k_nearest", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "traffic": NCU_TRAFFIC_NEAREST, "algorithmic_bytes_per_launch": algo, "ms": ms,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
            "workload": f"{Q} trees x {N} nodes (819 MB of float32 coordinates, > L2), one target each",
            "nodes_per_sec": Q * N / (ms * 1e-3)}


def mlp_roofline(peaks):
    """K4: actor MLP 733-1024-512-512-512-2 at B = 8192 on tcgen05 (bf16 x 3 split)."""
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
    algo = 3600384.0 * B
    peak = float(peaks.get("bf16_tflops", 1590.0))
    ach = algo / (ms * 1e-3) / 1e12
    big = torch.zeros(65536, ops.OBS_LD, dtype=torch.float32, device="cuda")
    big[:, :733] = torch.rand(65536, 733, device="cuda")
    ms_big = timed(lambda: ops.policy_forward(h, big), 5)
    ach_big = 3600384.0 * 65536 / (ms_big * 1e-3) / 1e12
    return {"kernel": "k_gemm_bf16x3 (+ split, head)", "bound": "tensor", "achieved": ach, "peak": peak,
            "unit": "TFLOP/s", "frac": ach / peak, "traffic": None, "ms": ms,
            "ncu": "profiles/r1j_gemm_ncu_summary.txt: tensor pipe 58.6 % active in the layer-0 GEMM",
            "at_65536_rows": {"achieved": ach_big, "frac": ach_big / peak, "ms": ms_big,
                              "kernel": "k_gemm_bf16x3_pair<256> (tcgen05.mma.cta_group::2, 256 x 256 tiles)",
                              "bf16_mma_tflops": 6.0 * ach_big * (1.835 - 0.786 * 0.4583) / 1.835},
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)" if "bf16_tflops" in peaks else "fallback",
            "note": "algorithmic float32 FLOP (3,600,384 per sample); the kernel issues 6 bf16 MMAs per "
                    "float32 product pair (bf16 x 3 operand split) to keep float32-level accuracy, so the "
                    "tensor pipe does 6x this work", "policy_samples_per_sec": B / (ms * 1e-3)}


def sub_rates(env, dwa, peaks):
    """SURVEY 8d sub-rates on the cfg4 map, each kernel alone on resident inputs (CUDA events):
    validity checks, clearances, local observations, estimator / classifier evaluations, DWA control
    calls (rollouts) per second, with their algorithmic work figures."""
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
        w = np.load(os.path.join(ROOT, "tests", "golden", "estimator_weights.npz"))
        est, cls = load_estimator({k[4:]: w[k] for k in w.files if k.startswith("est.")},
                                  {k[4:]: w[k] for k in w.files if k.startswith("cls.")})
        _, o32, _ = ops.local_obs(env.handle, states[:Bo], goals[:Bo], want64=False, want32=True)
        ms = timed(lambda: ops.estimator_forward(est.pair.handle, o32), 10)
        out["estimator_pairs_per_sec"] = Bo / (ms * 1e-3)
        out["estimator_tflops_algorithmic"] = 2 * 1.158038e6 * Bo / (ms * 1e-3) / 1e12     # 2 nets x 579,019 MAC x 2
    except Exception as e:                                                     # weights file not present
        out["estimator"] = f"skipped: {e}"
    # gym side: B environments per crlk_gym_step call (motion + _info + _reward + _obs)
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
    n_roll = float(r["n_rollouts"].double().sum().item()) if isinstance(r, dict) and "n_rollouts" in r else None
    out["dwa_controls_per_sec"] = Bd / (ms * 1e-3)
    if n_roll:
        out["rollouts_per_sec"] = n_roll / (ms * 1e-3)
        out["rollout_tflops_algorithmic"] = M_FLOP_ROLLOUT * n_roll / (ms * 1e-3) / 1e12
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_info()
    if args.workload == "cfg5":
        return run_cfg5(args)
    args.queries = args.queries or 1024
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from crl_kino_b200 import ops
    from crl_kino_b200.env import DifferentialDriveEnv
    from crl_kino_b200.planner.batched import BatchedExtender
    from crl_kino_b200.policy.dwa import DWA
    from crl_kino_b200.workloads import DENSE_DWA, dense_map

    rects = dense_map(n_cells=args.cells)
    env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi)
    env.set_obs(rects)
    dwa = DWA(env)
    dwa.set_dwa(**DENSE_DWA)
    if args.use_clearance:
        dwa.set_dwa(use_clearance=True)
        args.no_cpu = True          # the CPU restatement needs ~80 s per control step in this mode
        args.no_extra = True
    forest, batches = build_workload(args, env, rank)
    ext = BatchedExtender(env, forest, dwa=dwa)
    dev_batches = ops.as_dev(batches)
    K, W, Q = args.steps, args.warmup, args.queries
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ------------------------------------------------
    for w in range(W):
        ext.step(dev_batches[w])
    barrier()
    ext.stats.zero_()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    sampler = ClockSampler(local_rank)
    barrier()
    torch.cuda.profiler.start()          # ncu --profile-from-start off captures only the timed region
    t_wall = time.time()
    for k in range(K):
        flush.fill_(k & 0xff)                      # L2 flush between timed iterations (untimed)
        ext.step(dev_batches[W + k], events=evs[k])
    barrier()
    t_wall = time.time() - t_wall
    torch.cuda.profiler.stop()
    clocks = sampler.stop()
    t_near = sum(e[0].elapsed_time(e[1]) for e in evs)
    t_gath = sum(e[1].elapsed_time(e[2]) for e in evs)
    t_ext = sum(e[2].elapsed_time(e[3]) for e in evs)
    t_total_ms = t_near + t_gath + t_ext
    stats = ext.stats.cpu().numpy().astype(np.float64)
    n_ext, n_ctrl, n_roll = stats[0], stats[1], stats[2]
    status = ext.out["status"].cpu().numpy()

    # ---- end-to-end arm: host buffers through the public API -------------------
    # submit_host / collect: every step stages its samples in pinned memory, copies them to the
    # device, runs nearest + gather + extension and copies path / status / nodes back; two steps are
    # in flight so that the D2H of step k overlaps the kernels of step k + 1.
    for w in range(W):
        ext.collect(ext.submit_host(batches[w]))
    barrier()
    t0 = time.perf_counter()
    pending = None
    n_checked = 0
    for k in range(K):
        t = ext.submit_host(batches[W + k])
        if pending is not None:
            res = ext.collect(pending)
            n_checked += int(res["n_steps"].shape[0])
        pending = t
    res = ext.collect(pending)
    n_checked += int(res["n_steps"].shape[0])
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    assert n_checked == K * Q
    # the synchronous call (one step at a time), for comparison
    t0 = time.perf_counter()
    for k in range(K):
        ext.step_host(batches[W + k])
    t_e2e_sync = time.perf_counter() - t0

    plan = None
    if not args.no_extra:
        plan = planning_leg(args, env, dwa, rank)

    # ---- reduce over ranks (max time, summed work) ----------------------------------
    red = torch.tensor([t_total_ms, t_e2e, t_ext, plan["seconds"] if plan else 0.0,
                        plan["e2e_seconds"] if plan else 0.0], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([n_ext, n_ctrl, n_roll, plan["queries"] if plan else 0.0, plan["extensions"] if plan else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    t_total_ms, t_e2e, t_ext_max, t_plan, t_plan_e2e = [float(x) for x in red.cpu()]
    tot_ext, tot_ctrl, tot_roll, tot_q, tot_pext = [float(x) for x in cnt.cpu()]

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        sms = torch.cuda.get_device_properties(0).multi_processor_count
        fp64_peak = sms * FP64_LANES_PER_SM * 2 * sm_max * 1e6 / 1e12
        M = len(rects)
        flop = M_FLOP_ROLLOUT * n_roll + M_FLOP_PAIR * M * n_ctrl          # this rank's extend launches
        achieved = flop / (t_ext * 1e-3) / 1e12
        value = tot_ext / (t_total_ms * 1e-3)
        line = {
            "metric": "rrt_extensions_per_sec", "value": value, "unit": "extensions/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_total_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "cfg4: dense random-obstacle map, DWA steering",
                       "obstacles": M, "queries_per_gpu": Q, "tree_nodes": args.tree_nodes,
                       "rollouts_per_control_nominal": 4096,
                       "rollouts_per_control_realised": tot_roll / max(tot_ctrl, 1),
                       "control_steps_per_extension": tot_ctrl / max(tot_ext, 1),
                       "use_clearance": bool(args.use_clearance), "l2": "flushed between timed iterations (256 MiB write)",
                       "status_last_step": {"timeout": int((status == 0).sum()), "collision": int((status == 1).sum()),
                                            "reached": int((status == 2).sum())}},
            "e2e": {"value": tot_ext / t_e2e, "unit": "extensions/s",
                    "h2d_bytes_per_step": ext.h2d_bytes(), "d2h_bytes_per_step": ext.d2h_bytes(),
                    "api": "BatchedExtender.submit_host / collect (two steps in flight)",
                    "one_step_at_a_time_rank0": Q * K / t_e2e_sync},
            "gpu_launches": K * ext.launches_per_step,
            "kernel_ms_per_step": {"nearest": t_near / K, "gather": t_gath / K, "extend_dwa": t_ext / K},
            "roofline": {"kernel": "k_extend_dwa", "bound": "fp64", "achieved": achieved, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                         "traffic": NCU_TRAFFIC_EXTEND_DWA if (Q == 1024 and args.cells == 10000 and not args.use_clearance) else None,
                         "ncu": "profiles/r1t_extend_ncu_summary.txt: fp64 pipe 29.4 % while active, issue slots 51.7 %, 128 regs, 2 CTAs/SM; top stalls: fixed-latency dependencies of the float64 chains, then barriers",
                         "executed": {"fp64_pipe_pct_of_peak_while_active": 29.4, "issue_slots_pct": 51.7,
                                      "sm_throughput_pct_of_elapsed": 35.5, "warp_instructions_per_launch": 146798117,
                                      "source": "profiles/r1t_extend_ncu_summary.txt (ncu --set full, same command)"},
                         "peak_source": f"derived: {sms} SM x {FP64_LANES_PER_SM} FP64 FMA lanes x 2 x "
                                        f"{sm_max:.0f} MHz (MEASURED_PEAKS.json holds only HBM and bf16 peaks)",
                         "algorithmic_flop_per_launch": flop / K,
                         "note": "algorithmic count = 406/rollout + 860/(pose,obstacle) pair (SURVEY 8d); the kernel "
                                 "shares sincos chains across the grid and screens obstacles in FP32, so executed "
                                 "FLOP are far fewer"},
            "clocks": clocks, "wall_s_timed_region": t_wall,
        }
        if plan is not None:
            line["planning"] = {"queries_per_sec": tot_q / t_plan, "unit": "queries/s", "queries": int(tot_q),
                                "max_iter": plan["max_iter"], "ms": t_plan * 1e3,
                                "e2e": {"queries_per_sec": tot_q / t_plan_e2e, "h2d_bytes": plan["h2d_bytes"],
                                        "d2h_bytes": plan["d2h_bytes"]},
                                "reached_frac_rank0": plan["reached_frac"], "mean_nodes_rank0": plan["mean_nodes"],
                                "mean_samples_consumed_rank0": plan["mean_samples"],
                                "streams_exhausted_rank0": plan["streams_exhausted"],
                                "extensions_per_sec": tot_pext / t_plan, "gpu_launches": 1,
                                "device_sampling_rank0": {"queries_per_sec": plan["queries"] / plan["device_sampling_seconds"],
                                                          "reached_frac": plan["device_sampling_reached_frac"],
                                                          "note": "no sample streams: the kernel draws RRT.sample itself "
                                                                  "(counter-based generator), wall clock incl. the D2H of the results"},
                                "note": "Q full RRT.planning runs in one launch of k_plan_dwa (FusedRRT: one CTA per "
                                        "query from first sample to last node, persistent CTAs pull queries); trees "
                                        "and sample streams resident in HBM, CUDA-event time, best of 2 after warm-up; "
                                        "e2e adds the H2D of the pinned sample streams and the D2H of the results"}
            if "lockstep_seconds" in plan:
                line["planning"]["lockstep"] = {"queries_per_sec_rank0": plan["queries"] / plan["lockstep_seconds"],
                                                "rounds": plan["lockstep_rounds"],
                                                "identical_results": plan["lockstep_identical"],
                                                "note": "BatchedRRT: five launches per round, wall clock"}
            line["rooflines_other_kernels"] = [nearest_roofline(peaks), mlp_roofline(peaks)]
            line["sub_rates"] = sub_rates(env, dwa, peaks)
        if not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args, rects, forest, batches[W])
            if plan is not None:
                from oracle import oracle as orc
                cores = orc.cpu_count()
                line["planning"]["cpu_baseline"] = cpu_planning(rects, plan["inputs"], plan["max_iter"],
                                                                min(args.queries, cores), cores)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def run_cfg5(args):
    """cfg5: 65,536 independent queries sharded over the GPUs, 100,000-node trees, batched policy
    steering.  One step = nearest node of one sample per query + one RL extension (<= 50 policy
    steps) from that node toward the query's goal (the sample itself lies within the 1 m reach
    radius of such a dense tree, which would end every extension after one step)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_info()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from crl_kino_b200 import ops
    from crl_kino_b200.env import DifferentialDriveEnv, DifferentialDriveGym
    from crl_kino_b200.planner.batched import BatchedExtender, Forest
    from crl_kino_b200.planner.sharding import shard_range
    from crl_kino_b200.policy.rl_policy import load_policy
    from crl_kino_b200.workloads import dense_map, free_states

    N = 100000
    total_q = 65536
    lo, hi = shard_range(total_q, rank, world)
    Q = args.queries or (hi - lo)
    free_b, _ = torch.cuda.mem_get_info()
    fit = int(0.8 * free_b / (N * 64 + 50 * 5 * 8 + 3 * 736 * 4 * 2))
    capped = Q > fit
    Q = min(Q, fit)
    rects = dense_map(n_cells=args.cells)
    env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi)
    env.set_obs(rects)
    pol = load_policy(DifferentialDriveGym(), [1024, 512, 512, 512], None, seed=0)
    rng = np.random.default_rng(3 + rank)

    def clearance(c):
        return env.valid_state_batch(c, want_clearance=True)[1].cpu().numpy()
    pool = free_states(clearance, rng, 1 << 21, 0.0, batch=1 << 19)
    pool[:, 3:] = 0.0
    dpool = ops.as_dev(pool)
    forest = Forest(Q, N)
    g = torch.Generator(device="cuda").manual_seed(11 + rank)
    for q0 in range(0, Q, 128):
        q1 = min(Q, q0 + 128)
        idx = torch.randint(0, dpool.shape[0], (q1 - q0, N), generator=g, device="cuda")
        st = dpool[idx]
        forest.states[q0:q1] = st
        forest.x[q0:q1] = st[:, :, 0]; forest.y[q0:q1] = st[:, :, 1]
        forest.x32[q0:q1] = st[:, :, 0].float(); forest.y32[q0:q1] = st[:, :, 1].float()
        del st, idx
    forest.n_nodes[:] = N
    K, W = args.steps, args.warmup
    samples = [ops.as_dev(free_states(clearance, rng, Q, 0.0, moving=False)) for _ in range(K + W)]
    goals = ops.as_dev(free_states(clearance, rng, Q, 0.5))
    ext = BatchedExtender(env, forest, policy=pol, eps=0.0)
    ext.extend_targets = goals

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for w in range(W):
        ext.step(samples[w])
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    sampler = ClockSampler(local_rank)
    barrier()
    n_steps_tot = 0
    for k in range(K):
        out = ext.step(samples[W + k], events=evs[k])
    barrier()
    clocks = sampler.stop()
    t_near = sum(e[0].elapsed_time(e[1]) for e in evs)
    t_ext = sum(e[2].elapsed_time(e[3]) for e in evs)
    t_total = sum(e[0].elapsed_time(e[3]) for e in evs)
    mean_steps = float(out["n_steps"].double().mean().item())
    hs = [s.cpu().numpy() for s in samples]
    for w in range(W):
        ext.step_host(hs[w])
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        ext.step_host(hs[W + k])
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    red = torch.tensor([t_total, t_e2e, t_near, t_ext], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(Q * K)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    t_total, t_e2e, t_near, t_ext = [float(x) for x in red.cpu()]
    tot = float(cnt.item())
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        near_gbs = 8.0 * Q * N * K / (t_near * 1e-3) / 1e9
        print(json.dumps({
            "metric": "rrt_extensions_per_sec", "value": tot / (t_total * 1e-3), "unit": "extensions/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_total / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32 policy / f64 dynamics", "data": "synthetic",
            "config": {"workload": "cfg5: large trees, batched policy steering", "queries_total": int(tot / K),
                       "queries_per_gpu": Q, "capped_by_memory": bool(capped), "tree_nodes": N,
                       "obstacles": len(rects), "policy": "seeded random-init actor 733-1024-512-512-512-2, eps 0",
                       "mean_policy_steps_per_extension": mean_steps,
                       "l2": "per-step tree scan (>= 800 MB) exceeds L2"},
            "e2e": {"value": tot / t_e2e, "unit": "extensions/s", "h2d_bytes_per_step": ext.h2d_bytes(),
                    "d2h_bytes_per_step": ext.d2h_bytes()},
            "gpu_launches": K * ext.launches_per_step,
            "kernel_ms_per_step": {"nearest": t_near / K, "extend_rl": t_ext / K},
            "roofline": {"kernel": "k_nearest", "bound": "hbm", "achieved": near_gbs, "peak": hbm, "unit": "GB/s",
                         "frac": near_gbs / hbm, "traffic": None},
            "clocks": clocks}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_sample_run(rects, tree_states, samples, n, threads):
    """The oracle's C restatement of nearest + DWA extension on the first n queries."""
    from crl_kino_b200.workloads import DENSE_DWA
    from oracle import oracle as orc
    o = orc.Env()
    o.set_obs(rects)
    d = orc.Dwa(o, **DENSE_DWA)
    xs = np.ascontiguousarray(tree_states[:n, :, 0])
    ys = np.ascontiguousarray(tree_states[:n, :, 1])
    t0 = time.perf_counter()
    idx, _ = orc.nearest_batch(xs, ys, np.full(n, xs.shape[1], dtype=np.int32), samples[:n], threads=threads)
    froms = tree_states[np.arange(n), idx]
    r = d.extend_batch(froms, samples[:n], threads=threads)
    dt = time.perf_counter() - t0
    return n / dt, dt, int(r["n_steps"].sum())


def cpu_baseline(args, rects, forest, samples):
    from oracle import oracle as orc
    cores = orc.cpu_count()
    n = args.cpu_sample or min(args.queries, max(64 * cores, 256))
    ts = forest.states[:n].cpu().numpy()
    v, dt, steps = cpu_sample_run(rects, ts, samples, n, cores)
    return {"value": v, "unit": "extensions/s", "cores": cores, "kind": "port",
            "sample": f"first {n} queries of timed batch 0 (nearest + full DWA extension, {steps} control steps) "
                      f"through oracle/crlk_oracle.c on {cores} threads in {dt:.1f} s"}


def run_reference(args):
    args.queries = args.queries or 1024
    """--impl reference: the CPU restatement of the reference path (the Python
    reference itself cannot travel to the GPU box) on all host cores."""
    rank, _, world = dist_info()
    if rank != 0:
        return
    from crl_kino_b200.workloads import dense_map, random_states
    from oracle import oracle as orc
    rects = dense_map(n_cells=args.cells)
    o = orc.Env()
    o.set_obs(rects)
    cores = orc.cpu_count()
    rng = np.random.default_rng(100)
    n = args.cpu_sample or min(args.queries, max(16 * cores, 128))
    N = args.tree_nodes
    K, W = args.steps, args.warmup
    # Tree nodes: free states like the GPU arm's, but picked with a conservative raster test
    # (no obstacle within 1.5 m of the position, which implies clearance > 0.5 for the 0.86 m
    # circumradius) so that the setup does not spend minutes in the scalar clearance routine.
    def free(nn):
        res = 0.125
        gx = np.arange(-22.0, 22.0 + res, res)
        occ = np.zeros((len(gx), len(gx)), dtype=bool)
        pad = 1.5
        for l, b, r, t in rects:
            i0, i1 = np.searchsorted(gx, [l - pad, r + pad])
            j0, j1 = np.searchsorted(gx, [b - pad, t + pad])
            occ[max(i0 - 1, 0):i1 + 1, max(j0 - 1, 0):j1 + 1] = True
        out = []
        while sum(len(x) for x in out) < nn:
            c = random_states(rng, 1 << 16)
            ii = np.clip(np.searchsorted(gx, c[:, 0]) - 1, 0, len(gx) - 1)
            jj = np.clip(np.searchsorted(gx, c[:, 1]) - 1, 0, len(gx) - 1)
            out.append(c[~occ[ii, jj]])
        return np.concatenate(out)[:nn]
    nodes = free(n * N)
    nodes[:, 3:] = 0
    trees = nodes.reshape(n, N, 5)
    xs, ys = np.ascontiguousarray(trees[:, :, 0]), np.ascontiguousarray(trees[:, :, 1])
    batches = []
    for b in range(K + W):
        s = np.zeros((n, 5)); need = np.ones(n, dtype=bool)
        while need.any():
            c = random_states(rng, n, moving=True)
            v = o.valid_state_batch(c, clearance=False)[0]
            _, d = orc.nearest_batch(xs, ys, np.full(n, N, dtype=np.int32), c)
            ok = need & v & (d >= 1.0) & (d < 20)
            s[ok] = c[ok]; need[ok] = False
        batches.append(s)
    for w in range(min(W, 1)):
        cpu_sample_run(rects, trees, batches[w], n, cores)
    t0 = time.perf_counter()
    tot = 0
    for k in range(K):
        cpu_sample_run(rects, trees, batches[W + k], n, cores)
        tot += n
    dt = time.perf_counter() - t0
    v = tot / dt
    print(json.dumps({
        "impl": "reference", "metric": "rrt_extensions_per_sec", "value": v, "unit": "extensions/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg4: dense random-obstacle map, DWA steering", "obstacles": len(rects),
                   "queries_per_step": n, "tree_nodes": N},
        "cpu_baseline": {"value": v, "unit": "extensions/s", "cores": cores, "kind": "port",
                         "sample": f"{n} queries per step (bounded sample of the 1024-query batch), nearest + full "
                                   f"DWA extension through oracle/crlk_oracle.c on {cores} threads"},
        "e2e": {"value": v, "unit": "extensions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
