"""Actor MLP micro-benchmark: tensor-core (tcgen05 bf16x3) vs SIMT float32 path.
Run once per path: CRLK_MLP_SIMT=1 selects the SIMT kernels."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crl_kino_b200 import ops
from crl_kino_b200.env import DifferentialDriveGym
from crl_kino_b200.policy.rl_policy import load_policy

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
pol = load_policy(DifferentialDriveGym(), [1024, 512, 512, 512], None, seed=0)
rng = np.random.default_rng(0)
obs = np.concatenate([rng.normal(0, 5, (B, 4)), (rng.uniform(0, 1, (B, 729)) > 0.3).astype(float)], 1)
o32 = ops.pad_obs(obs)
h = pol.actor.handle
act = ops.policy_forward(h, o32)
# float64 reference of the same float32 weights
x = obs.astype(np.float32).astype(np.float64)
for i, (W, b) in enumerate(zip(pol.actor.weights, pol.actor.biases)):
    x = x @ W.astype(np.float64).T + b.astype(np.float64)
    if i < 4: x = np.maximum(x, 0)
lo, hi = pol.actor.low.astype(np.float64), pol.actor.high.astype(np.float64)
ref = np.clip(np.tanh(x) * (hi - lo) / 2 + (lo + hi) / 2, lo, hi)
err = np.abs(act.cpu().numpy() - ref)
for _ in range(3): ops.policy_forward(h, o32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n): ops.policy_forward(h, o32)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
flop = 3600384.0 * B
print(json.dumps({"path": "simt" if os.environ.get("CRLK_MLP_SIMT") else "tcgen05_bf16x3", "B": B, "ms": ms,
                  "algorithmic_tflops": flop / ms / 1e9, "max_abs_err_vs_f64": float(err.max()),
                  "mean_abs_err": float(err.mean())}))
