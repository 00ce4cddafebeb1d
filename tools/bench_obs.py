"""k_local_obs alone on the cfg4 map (for ncu): float32 rows, 8192 observations."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crl_kino_b200 import ops
from crl_kino_b200.env import DifferentialDriveEnv
from crl_kino_b200.workloads import dense_map, free_states
env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi); env.set_obs(dense_map())
rng = np.random.default_rng(0)
clr = lambda c: env.valid_state_batch(c, want_clearance=True)[1].cpu().numpy()
B = 8192
s = ops.as_dev(free_states(clr, rng, B, 0.0, moving=True)); g = ops.as_dev(free_states(clr, rng, B, 0.5))
for want_near in (False, True):
    f = lambda: ops.local_obs(env.handle, s, g, want64=False, want32=True, want_near=want_near)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    print(f"want_near={want_near}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us for {B} observations")
