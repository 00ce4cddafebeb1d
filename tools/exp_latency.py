"""Extend kernel: per-control-step latency (one CTA alone) vs loaded throughput."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crl_kino_b200 import ops
from crl_kino_b200.env import DifferentialDriveEnv
from crl_kino_b200.policy.dwa import DWA
from crl_kino_b200.workloads import DENSE_DWA, dense_map, free_states, random_states
rects = dense_map(); env = DifferentialDriveEnv(1.0,-0.1,np.pi,1.0,np.pi); env.set_obs(rects)
dwa = DWA(env); dwa.set_dwa(**DENSE_DWA); dp = dwa.params(); xp = ops.extend_params()
rng = np.random.default_rng(0)
clr = lambda c: env.valid_state_batch(c, want_clearance=True)[1].cpu().numpy()
cand = free_states(clr, rng, 40000, 2.5); cand[:, 3:] = 0
tg = cand.copy(); ang = rng.uniform(-np.pi, np.pi, len(cand)); tg[:, 0] += 15*np.cos(ang); tg[:, 1] += 15*np.sin(ang)
out = ops.extend_dwa(env.handle, dp, xp, cand, tg)
ns = out["n_steps"].cpu().numpy(); order = np.argsort(-ns)
print("steps of the longest:", ns[order[:5]], "n with 50:", (ns == 50).sum())
sel = order[:4096]
def t(Q, n=5):
    f, g = ops.as_dev(cand[sel[:Q]]), ops.as_dev(tg[sel[:Q]])
    steps = ns[sel[:Q]].sum()
    for _ in range(2): ops.extend_dwa(env.handle, dp, xp, f, g)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): ops.extend_dwa(env.handle, dp, xp, f, g)
    e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)/n
    print(f"Q={Q:5d} steps={steps:7d} ms={ms:8.3f} us/step(chain)={1e3*ms/ns[sel[:Q]].max():7.2f} us/step/SM(throughput)={1e3*ms*148/steps:7.2f}")
for Q in (1, 148, 296, 444, 888, 1776, 4096): t(Q)
