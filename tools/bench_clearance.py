"""k_clearance alone on the cfg4 map (for ncu): 131072 random poses."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crl_kino_b200 import ops
from crl_kino_b200.env import DifferentialDriveEnv
from crl_kino_b200.workloads import dense_map, random_states
env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi); env.set_obs(dense_map())
rng = np.random.default_rng(17)
B = 1 << 17
s = ops.as_dev(random_states(rng, B, moving=True))
f = lambda: ops.valid_state(env.handle, s, want_clearance=True)
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
v, c, _ = f()
c = c.cpu().numpy()
print(f"{ms:.3f} ms for {B} clearances ({B / ms * 1e3:.3e}/s); colliding {np.mean(c < 0):.3f}, mean clearance of the free ones {c[c >= 0].mean():.3f}, max {c.max():.2f}")
