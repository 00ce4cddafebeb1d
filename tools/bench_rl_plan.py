"""The learned-steering planning leg of bench.py alone (RRT_RL / RRT_RL_Estimator, lock-step BatchedRRT).
usage: python tools/bench_rl_plan.py [queries] [max_iter] [--cpu]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from crl_kino_b200.env import DifferentialDriveEnv
from crl_kino_b200.workloads import dense_map

class A: pass
a = A()
a.plan_rl_queries = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
a.plan_rl_iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
rects = dense_map()
env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi)
env.set_obs(rects)
print(json.dumps(bench.planning_rl_leg(a, env, rects, "--cpu" not in sys.argv), indent=1))
