#!/usr/bin/env python
"""The reference's examples/test.py cases on the B200 path (no plotting).

    python examples/test.py --case rrt                # DWA-steered kinodynamic RRT, one query (BASELINE config 1)
    python examples/test.py --case rl_rrt             # policy steering (config 2; seeded random-init actor: the
                                                      #   reference's trained policy.pth is not in its checkout)
    python examples/test.py --case rl_rrt_estimator   # learned parent selection + policy steering
    python examples/test.py --case gym                # DifferentialDriveGym episode driven by the policy
    python examples/test.py --case batch --queries 64 # many RRT queries in one launch (FusedRRT)

Maps come from --data DIR (the reference's data/obstacles/*.pkl, read as they are) or, by default,
from the copies of test_env1 / test_env2 kept in data/fixtures.npz.  Trees and paths are
written as pickles the reference's own scripts can load (--out DIR).
"""
import argparse
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from crl_kino_b200.env import DifferentialDriveEnv, DifferentialDriveGym  # noqa: E402
from crl_kino_b200.utils import io as cio  # noqa: E402


def load_map(args, name):
    if args.data:
        return cio.load_obstacles(os.path.join(args.data, name + ".pkl"))
    fx = np.load(os.path.join(ROOT, "data", "fixtures.npz"))
    return fx["map_" + name]


def make_policy(seed=0):
    from crl_kino_b200.policy.rl_policy import load_policy
    path = os.environ.get("CRLK_POLICY_PTH")              # a policy.pth of the reference, if you have one
    return load_policy(DifferentialDriveGym(), [1024, 512, 512, 512], path, seed=seed)


def make_estimator():
    from crl_kino_b200.estimator import load_estimator
    w = np.load(os.path.join(ROOT, "data", "estimator_weights.npz"))
    return load_estimator({k[4:]: w[k] for k in w.files if k.startswith("est.")},
                          {k[4:]: w[k] for k in w.files if k.startswith("cls.")})


def report(planner, path, args, tag):
    print(f"{tag}: reached={planner.reach_exactly} nodes={len(planner.node_list)} "
          f"TTR={0.2 * (len(path) - 1):.1f}s planning_time={planner.planning_time:.3f}s")
    if args.out:
        os.makedirs(args.out, exist_ok=True)
        cio.save_tree(planner.node_list, os.path.join(args.out, tag + "_tree.pkl"))
        cio.save_path(path, os.path.join(args.out, tag + "_path.pkl"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="rrt", choices=["rrt", "rl_rrt", "rl_rrt_estimator", "gym", "batch"])
    ap.add_argument("--env", default="test_env1")
    ap.add_argument("--data", default=None, help="directory with the reference's obstacle pickles")
    ap.add_argument("--out", default=None, help="write tree / path pickles here")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-iter", type=int, default=500)
    ap.add_argument("--queries", type=int, default=64)
    args = ap.parse_args()
    random.seed(args.seed)
    np.random.seed(args.seed)
    start = np.array([-5, -15, 0, 0, 0.0])
    goal = np.array([10, 10, 0, 0, 0.0])
    env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi)
    env.set_obs(load_map(args, args.env))

    if args.case == "rrt":
        from crl_kino_b200.planner.rrt import RRT
        planner = RRT(env, max_iter=args.max_iter)
    elif args.case == "rl_rrt":
        from crl_kino_b200.planner.rrt_rl import RRT_RL
        planner = RRT_RL(env, make_policy(args.seed), max_iter=args.max_iter)
    elif args.case == "rl_rrt_estimator":
        from crl_kino_b200.planner.rrt_rl_estimator import RRT_RL_Estimator
        est, cls = make_estimator()
        planner = RRT_RL_Estimator(env, make_policy(args.seed), est, cls, max_iter=args.max_iter)
    elif args.case == "gym":
        from crl_kino_b200.policy.rl_policy import policy_forward
        gym_env = DifferentialDriveGym(env, obs_list_list=[load_map(args, args.env)])
        policy = make_policy(args.seed)
        obs = gym_env.reset()
        total = 0.0
        for t in range(200):
            action = policy_forward(policy, obs, eps=0.05)
            obs, reward, done, info = gym_env.step(action[0])
            total += reward
            if done:
                break
        print(f"gym: {t + 1} steps, return {total:.2f}, goal={info['goal']} collision={info['collision']}")
        return
    else:
        from crl_kino_b200.planner.batched import FusedRRT
        from crl_kino_b200.workloads import query_pairs, sample_stream
        Q = args.queries
        starts, goals = query_pairs(lambda c: env.valid_state_batch(c, want_clearance=True)[1].cpu().numpy(), Q,
                                    seed=args.seed)
        sb = env.get_bounds()["state_bounds"]
        iters = min(args.max_iter, 100)
        streams = np.stack([sample_stream(sb, goals[q], 6 * iters, seed=1000 * args.seed + q) for q in range(Q)])
        p = FusedRRT(env, starts, goals, max_iter=iters)
        t0 = time.time()
        reached = p.plan(streams)
        dt = time.time() - t0
        print(f"batch: {Q} queries in {dt * 1e3:.1f} ms ({Q / dt:.0f} queries/s), reached {reached.mean():.2f}, "
              f"mean nodes {p.forest.n_nodes.double().mean().item():.1f}")
        return

    planner.verbose = False
    planner.set_start_and_goal(start, goal)
    path = planner.planning()
    report(planner, path, args, args.case)


if __name__ == "__main__":
    main()
