"""One crlk_extend_rl pass (cfg5 shape, Q queries) inside a cudaProfiler range, for
`ncu --profile-from-start off`; prints the CUDA-event time of an unprofiled pass first."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crl_kino_b200 import ops
from crl_kino_b200.env import DifferentialDriveEnv, DifferentialDriveGym
from crl_kino_b200.policy.rl_policy import load_policy
from crl_kino_b200.workloads import dense_map, free_states

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi); env.set_obs(dense_map())
pol = load_policy(DifferentialDriveGym(), [1024, 512, 512, 512], None, seed=0)
rng = np.random.default_rng(0)
clr = lambda c: env.valid_state_batch(c, want_clearance=True)[1].cpu().numpy()
froms = free_states(clr, rng, Q, 0.0); froms[:, 3:] = 0
targets = free_states(clr, rng, Q, 0.5)
fs, tg = ops.as_dev(froms), ops.as_dev(targets)
xp = ops.extend_params()
out = None
for _ in range(2):
    out = ops.extend_rl(env.handle, pol.actor.handle, xp, fs, tg, None, 0.0, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    out = ops.extend_rl(env.handle, pol.actor.handle, xp, fs, tg, None, 0.0, out=out)
e1.record(); torch.cuda.synchronize()
ns = out["n_steps"].cpu().numpy()
print(json.dumps({"Q": Q, "extend_rl_ms": e0.elapsed_time(e1) / 3, "mean_steps": float(ns.mean()),
                  "live_rows_by_step": [int((ns >= k).sum()) for k in (1, 5, 10, 20, 30, 40, 50)]}))
torch.cuda.profiler.start()
out = ops.extend_rl(env.handle, pol.actor.handle, xp, fs, tg, None, 0.0, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
