"""Estimator / classifier CNN pair alone (for ncu): 16384 observations on the cfg4 map."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crl_kino_b200 import ops
from crl_kino_b200.env import DifferentialDriveEnv
from crl_kino_b200.estimator import load_estimator
from crl_kino_b200.workloads import dense_map, random_states
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi); env.set_obs(dense_map())
w = np.load(os.path.join(ROOT, "tests", "golden", "estimator_weights.npz"))
est, cls = load_estimator({k[4:]: w[k] for k in w.files if k.startswith("est.")},
                          {k[4:]: w[k] for k in w.files if k.startswith("cls.")})
rng = np.random.default_rng(0)
B = 16384
_, o32, _ = ops.local_obs(env.handle, random_states(rng, B, moving=True), random_states(rng, B), want64=False, want32=True)
for _ in range(3):
    ops.estimator_forward(est.pair.handle, o32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.estimator_forward(est.pair.handle, o32)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"B={B} {ms:.3f} ms  {B / ms * 1e3:.3e} pairs/s  {2 * 1.158038e6 * B / ms * 1e3 / 1e12:.2f} TFLOP/s (algorithmic)")
