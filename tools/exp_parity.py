#!/usr/bin/env python
"""Experiment: how often does the DWA argmin of k_extend_dwa differ from the oracle's, and why?
cfg4 map, Q extensions from free states; prints flag counts, diverged extensions and, for the
first control step at which the chosen rollout differs, the oracle's costs of both candidates."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    which = sys.argv[2] if len(sys.argv) > 2 else "dense"
    import torch
    from crl_kino_b200 import ops
    from crl_kino_b200.env import DifferentialDriveEnv
    from crl_kino_b200.policy.dwa import DWA
    from crl_kino_b200.workloads import DENSE_DWA, dense_map, query_pairs
    from oracle import oracle as orc
    rects = dense_map()
    kw = DENSE_DWA if which == "dense" else dict(dt=0.2, v_res=0.02)
    env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi)
    env.set_obs(rects)
    o = orc.Env()
    o.set_obs(rects)

    def clearance(c):
        return env.valid_state_batch(c, want_clearance=True)[1].cpu().numpy()
    s, g = query_pairs(clearance, Q)
    d = DWA(env)
    d.set_dwa(**kw)
    out = ops.extend_dwa(env.handle, d.params(), ops.extend_params(), s, g, want_ctrl=True)
    torch.cuda.synchronize()
    t0 = time.time()
    od = orc.Dwa(o, **kw)
    ref = od.extend_batch(s, g, want_ctrl=True)
    print(f"oracle: {Q} extensions in {time.time() - t0:.1f}s on {orc.cpu_count()} threads")
    fl = out["flags"].cpu().numpy()
    ns = out["n_steps"].cpu().numpy()
    st = out["status"].cpu().numpy()
    path = out["path"].cpu().numpy()
    ci = out["ctrl_idx"].cpu().numpy()
    res = dict(Q=Q, near_boundary=int((fl & 1).astype(bool).sum()), near_tie=int((fl & 2).astype(bool).sum()),
               clean=int((fl == 0).sum()))
    idx_diff, state_diff, shape_diff = [], [], []
    for q in range(Q):
        n = ref["n_steps"][q]
        if ns[q] != n or st[q] != ref["status"][q]:
            shape_diff.append(q)
        m = min(n, ns[q])
        if not np.array_equal(ci[q, :m], ref["ctrl_idx"][q, :m]):
            idx_diff.append(q)
        if not np.allclose(path[q, :m], ref["path"][q, :m], rtol=1e-9, atol=1e-11):
            state_diff.append(q)
    res.update(idx_diff=len(idx_diff), state_diff=len(state_diff), shape_diff=len(shape_diff),
               idx_diff_unflagged=int(sum(1 for q in idx_diff if not (fl[q] & 2))),
               state_diff_unflagged=int(sum(1 for q in state_diff if fl[q] == 0)),
               shape_diff_unflagged=int(sum(1 for q in shape_diff if fl[q] == 0)))
    print(json.dumps(res))
    # anatomy of the first few index differences
    for q in idx_diff[:12]:
        n = min(ref["n_steps"][q], ns[q])
        k = int(np.argmax(ci[q, :n] != ref["ctrl_idx"][q, :n]))
        state = s[q] if k == 0 else ref["path"][q, k - 1]
        same_state = k == 0 or np.array_equal(path[q, k - 1], ref["path"][q, k - 1])
        r = od.control(state, g[q], want_costs=True)
        a, b = int(ci[q, k]), int(ref["ctrl_idx"][q, k])
        cs = r["cost_sum"]
        dwn = o.dynamic_window(state, kw["dt"])
        nw = len(orc.arange(dwn[1, 0], dwn[1, 1] + d.w_res, d.w_res))
        print(f"q{q} step{k} flags={fl[q]} same_state_before={same_state} gpu_idx={a} (i={a // nw}, j={a % nw}) "
              f"oracle_idx={b} (i={b // nw}, j={b % nw}) R={r['R']} nw={nw} cost_gpu_choice={cs[a]!r} cost_oracle={cs[b]!r} "
              f"rel_gap={(cs[a] - cs[b]) / cs[b]:.3e} v={state[3]!r} w={state[4]!r}")
    # control level: one DWA control call per state, GPU vs oracle
    rng = np.random.default_rng(5)
    from crl_kino_b200.workloads import random_states
    B = 2048
    stt = random_states(rng, B, moving=True)
    stt[: B // 2, 3:] = 0.0
    gl = random_states(rng, B)
    r = ops.dwa_control(env.handle, d.params(), stt, gl)
    rb = od.control_batch(stt, gl)
    gi = r["opt_idx"].cpu().numpy()
    tie = r["near_tie"].cpu().numpy().astype(bool)
    diff = gi != rb["opt_idx"]
    print(json.dumps(dict(control_calls=B, near_tie=int(tie.sum()), idx_diff=int(diff.sum()),
                          idx_diff_unflagged=int((diff & ~tie).sum()))))


if __name__ == "__main__":
    main()
