"""Per-phase cycle breakdown of the control step (build with CRLK_DEFS=CRLK_PHASE_TIMING)."""
import os, sys, ctypes, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crl_kino_b200 import ops
from crl_kino_b200._lib import lib
from crl_kino_b200.env import DifferentialDriveEnv
from crl_kino_b200.policy.dwa import DWA
from crl_kino_b200.workloads import DENSE_DWA, dense_map, free_states
rects = dense_map(); env = DifferentialDriveEnv(1.0,-0.1,np.pi,1.0,np.pi); env.set_obs(rects)
dwa = DWA(env); dwa.set_dwa(**DENSE_DWA)
if os.environ.get('USE_CLEARANCE'): dwa.set_dwa(use_clearance=True)
dp = dwa.params(); xp = ops.extend_params()
rng = np.random.default_rng(0)
clr = lambda c: env.valid_state_batch(c, want_clearance=True)[1].cpu().numpy()
cand = free_states(clr, rng, 4000, 2.5); cand[:, 3:] = 0
tg = cand.copy(); ang = rng.uniform(-np.pi, np.pi, len(cand)); tg[:, 0] += 15*np.cos(ang); tg[:, 1] += 15*np.sin(ang)
names = ["wait-prev", "tables", "rollouts", "costs+argmin", "f:merge+tie", "column sums", "f:row1", "f:closing-barrier", "collision", "f:epilogue"]
raw = ctypes.CDLL(lib()._name)
for Q in ((1, 296) if os.environ.get('USE_CLEARANCE') else tuple(int(v) for v in os.environ.get('PHASE_Q', '1,444,1776').split(','))):
    f, g = ops.as_dev(cand[:Q]), ops.as_dev(tg[:Q])
    buf = (ctypes.c_ulonglong * 16)()
    ops.extend_dwa(env.handle, dp, xp, f, g); torch.cuda.synchronize()
    raw.crlk_debug_phase_cycles(buf, 1)
    out = ops.extend_dwa(env.handle, dp, xp, f, g); torch.cuda.synchronize()
    raw.crlk_debug_phase_cycles(buf, 1)
    steps = int(out["n_steps"].sum().item())
    tot = sum(buf[:10])
    print(f"Q={Q} control steps={steps} cycles/step={tot/steps:.0f} ({tot/steps/1.965e3:.1f} us)")
    print("   " + "  ".join(f"{n}={buf[i]/steps:.0f}" for i, n in enumerate(names)))
    if os.environ.get('USE_CLEARANCE'):
        print(f"   obstacle list: mean length {buf[10]/steps:.0f}, overflow in {100*buf[11]/steps:.1f}% of the control steps")
