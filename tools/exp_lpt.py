import sys, numpy as np, torch, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from crl_kino_b200 import ops
from crl_kino_b200.env import DifferentialDriveEnv
from crl_kino_b200.planner.batched import BatchedExtender
from crl_kino_b200.policy.dwa import DWA
from crl_kino_b200.workloads import DENSE_DWA, dense_map
class A: pass
a = A(); a.queries=1024; a.tree_nodes=1024; a.steps=6; a.warmup=2; a.cells=10000
rects = dense_map(); env = DifferentialDriveEnv(1.0,-0.1,np.pi,1.0,np.pi); env.set_obs(rects)
dwa = DWA(env); dwa.set_dwa(**DENSE_DWA)
forest, batches = bench.build_workload(a, env, 0)
ext = BatchedExtender(env, forest, dwa=dwa)
db = ops.as_dev(batches)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
out = ext.step(db[0]); ns = out['n_steps'].cpu().numpy(); d = out['nearest_dist'].cpu().numpy()
print('n_steps: mean %.2f max %d p99 %d; hist' % (ns.mean(), ns.max(), np.percentile(ns,99)), np.bincount(ns)[:52])
print('corr(dist, steps)', np.corrcoef(d, ns)[0,1])
fs = ext.from_states.clone(); tg = db[0]
xp = ext.xp; dp = dwa.params()
t0 = timeit(lambda: ops.extend_dwa(env.handle, dp, xp, fs, tg))
order = torch.argsort(out['nearest_dist'], descending=True)
fs2, tg2 = fs[order].contiguous(), tg[order].contiguous()
t1 = timeit(lambda: ops.extend_dwa(env.handle, dp, xp, fs2, tg2))
order = torch.argsort(out['n_steps'], descending=True)
fs3, tg3 = fs[order].contiguous(), tg[order].contiguous()
t2 = timeit(lambda: ops.extend_dwa(env.handle, dp, xp, fs3, tg3))
print('extend ms: index order %.3f, by dist desc %.3f, oracle LPT (true steps) %.3f' % (t0, t1, t2))
print('sum steps', ns.sum(), 'ideal per-CTA steps', ns.sum()/444)
