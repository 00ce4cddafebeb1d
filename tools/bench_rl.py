"""RL-steering extension micro-benchmark (cfg5 shape on one GPU): per-kernel times."""
import os, sys, json, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crl_kino_b200 import ops
from crl_kino_b200.env import DifferentialDriveEnv, DifferentialDriveGym
from crl_kino_b200.policy.rl_policy import load_policy
from crl_kino_b200.workloads import dense_map, free_states, random_states

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rects = dense_map()
env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi); env.set_obs(rects)
pol = load_policy(DifferentialDriveGym(), [1024, 512, 512, 512], None, seed=0)
rng = np.random.default_rng(0)
clr = lambda c: env.valid_state_batch(c, want_clearance=True)[1].cpu().numpy()
froms = free_states(clr, rng, Q, 0.5); froms[:, 3:] = 0
targets = random_states(rng, Q)
fs, tg = ops.as_dev(froms), ops.as_dev(targets)
xp = ops.extend_params()
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
out = ops.extend_rl(env.handle, pol.actor.handle, xp, fs, tg, None, 0.0)
ns = out["n_steps"].cpu().numpy(); st = out["status"].cpu().numpy()
res = {"Q": Q, "mean_steps": float(ns.mean()), "status": np.bincount(st, minlength=3).tolist()}
res["extend_rl_ms"] = t(lambda: ops.extend_rl(env.handle, pol.actor.handle, xp, fs, tg, None, 0.0), 3)
res["local_obs_ms"] = t(lambda: ops.local_obs(env.handle, fs, tg, want64=False, want32=True))
o32 = ops.local_obs(env.handle, fs, tg, want64=False, want32=True)[1]
res["policy_forward_ms"] = t(lambda: ops.policy_forward(pol.actor.handle, o32))
res["valid_state_ms"] = t(lambda: ops.valid_state(env.handle, fs))
res["extensions_per_sec"] = Q / (res["extend_rl_ms"] * 1e-3)
print(json.dumps(res))
