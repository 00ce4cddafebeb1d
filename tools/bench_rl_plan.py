"""RL-steered planning (RRT_RL, lock-step BatchedRRT) throughput on the cfg4 map."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crl_kino_b200 import ops
from crl_kino_b200.env import DifferentialDriveEnv, DifferentialDriveGym
from crl_kino_b200.planner.batched import BatchedRRT
from crl_kino_b200.policy.rl_policy import load_policy
from crl_kino_b200.workloads import dense_map, query_pairs
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi); env.set_obs(dense_map())
pol = load_policy(DifferentialDriveGym(), [1024, 512, 512, 512], None, seed=0)
clr = lambda c: env.valid_state_batch(c, want_clearance=True)[1].cpu().numpy()
starts, goals = query_pairs(clr, Q, seed=7)
sb = env.get_bounds()["state_bounds"]
S = iters * 4 + 16
rng = np.random.default_rng(300)
streams = rng.uniform(sb[:, 0], sb[:, 1], (Q, S, 5))
pick = rng.integers(0, 101, (Q, S)) <= 5
streams[pick] = np.repeat(goals[:, None, :], S, 1)[pick]
d = torch.as_tensor(streams).cuda()
for rep in range(2):
    p = BatchedRRT(env, starts, goals, max_iter=iters, policy=pol, eps=0.0, keep_paths=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    reached = p.plan(d, check_every=8)
    dt = time.perf_counter() - t0
    print(f"Q={Q} max_iter={iters}: {dt*1e3:.1f} ms, {Q/dt:.0f} queries/s, rounds {p.samples_used}, reached {reached.mean():.3f}, "
          f"mean nodes {p.forest.n_nodes.double().mean().item():.1f}")
