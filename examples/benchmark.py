#!/usr/bin/env python
"""The reference's examples/benchmark.py on the B200 path: every ordered (start, goal) pair of the
benchmark positions on test_env1 / test_env2 (BASELINE config 3), results written in the
reference's planning_results format ({'runtime': [...], 'path_len': [...]}, runtime -1 = goal not
reached).

    python examples/benchmark.py --planner rrt              # DWA steering: all 12 pairs in ONE launch
    python examples/benchmark.py --planner rl_rrt_estimator # the planner the reference's script runs
"""
import argparse
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))

from crl_kino_b200.env import DifferentialDriveEnv  # noqa: E402
from crl_kino_b200.utils import io as cio  # noqa: E402


def inputs(args, name):
    if args.data:
        obs = cio.load_obstacles(os.path.join(args.data, "data", "obstacles", f"test_{name}.pkl"))
        pos = cio.load_positions(os.path.join(args.data, "examples", f"benchmark_position_test_{name}.pkl"))
        return obs, pos
    fx = np.load(os.path.join(ROOT, "data", "fixtures.npz"))
    return fx["map_test_" + name], fx["positions_" + name]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--planner", default="rrt", choices=["rrt", "rl_rrt", "rl_rrt_estimator"])
    ap.add_argument("--data", default=None, help="root of a reference checkout (for its pickles)")
    ap.add_argument("--out", default=".")
    ap.add_argument("--max-iter", type=int, default=500)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    random.seed(args.seed)
    np.random.seed(args.seed)
    for name in ("env1", "env2"):
        obs, positions = inputs(args, name)
        env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi)
        env.set_obs(obs)
        pairs = cio.benchmark_pairs(positions)
        if args.planner == "rrt":
            from crl_kino_b200.planner.batched import FusedRRT
            from crl_kino_b200.workloads import sample_stream
            starts = np.array([positions[i] for i, _ in pairs])
            goals = np.array([positions[j] for _, j in pairs])
            sb = env.get_bounds()["state_bounds"]
            streams = np.stack([sample_stream(sb, goals[q], 8 * args.max_iter, seed=100 * args.seed + q)
                                for q in range(len(pairs))])
            p = FusedRRT(env, starts, goals, max_iter=args.max_iter)
            t0 = time.time()
            reached = p.plan(streams)
            dt = time.time() - t0
            lens = [len(p.final_course(q)) for q in range(len(pairs))]
            data = cio.planning_results(reached, [dt / len(pairs)] * len(pairs), lens)
        else:
            from test import make_estimator, make_policy
            if args.planner == "rl_rrt":
                from crl_kino_b200.planner.rrt_rl import RRT_RL
                planner = RRT_RL(env, make_policy(args.seed), max_iter=args.max_iter)
            else:
                from crl_kino_b200.planner.rrt_rl_estimator import RRT_RL_Estimator
                est, cls = make_estimator()
                planner = RRT_RL_Estimator(env, make_policy(args.seed), est, cls, max_iter=args.max_iter)
            planner.verbose = False
            ok, times, lens = [], [], []
            for i, j in pairs:
                planner.set_start_and_goal(positions[i], positions[j])
                path = planner.planning()
                ok.append(planner.reach_exactly); times.append(planner.planning_time); lens.append(len(path))
            data = cio.planning_results(ok, times, lens)
        fname = os.path.join(args.out, f"planning_results_{args.planner}_test_{name}")
        cio.save_results(data, fname)
        print(f"test_{name}: {int((data['runtime'] >= 0).sum())}/{len(pairs)} reached, "
              f"mean runtime {np.mean(data['runtime'][data['runtime'] >= 0]) if (data['runtime'] >= 0).any() else float('nan'):.3f}s"
              f" -> {fname}.pkl")


if __name__ == "__main__":
    main()
