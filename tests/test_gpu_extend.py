"""Fused extend kernel and the RRT drop-in class vs goldens / oracle."""
import random

import numpy as np
import pytest

from conftest import golden
from gpu_util import fixture_map, make_envs, np_
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _extend(env, froms, targets, **dwa_kw):
    from crl_kino_b200 import ops
    from crl_kino_b200.policy.dwa import DWA
    d = DWA(env)
    d.set_dwa(**(dwa_kw or dict(dt=0.2, v_res=0.02)))
    return ops.extend_dwa(env.handle, d.params(), ops.extend_params(), froms, targets, want_ctrl=True)


def _compare(out, ref, what, max_flagged=0.15, max_diverged=0):
    """The contract of include/crlk.h: step counts, status, emitted nodes and the chosen rollout of every
    control step exactly the oracle's, executed states to 1e-12 (contract: 1e-4) -- unless the kernel itself
    reported the extension: flags bit 0 (a collision decision within 1e-6 of the threshold) or bit 1 (an
    argmin decision that one misrounded last bit of the C library's atan2 could change).  Reported
    extensions may differ, at most `max_diverged` of them do, and at most `max_flagged` of the batch is
    reported at all.  Executed states are the oracle's BITS except where the C library misrounds a sin /
    cos (6e-4 of its results, one ulp, which survives into a coordinate about once in 10^4 steps): at
    least 98 % of the extensions must be bit-identical."""
    fl = np_(out["flags"])
    assert (fl != 0).mean() <= max_flagged, f"{what}: {(fl != 0).sum()} of {len(fl)} extensions reported"
    path, rpath = np_(out["path"]), ref["path"]
    diverged, bitwise = 0, 0
    for q in range(len(fl)):
        n = int(ref["n_steps"][q])
        bitwise += int(np_(out["n_steps"])[q] == n and np.array_equal(path[q, :n], rpath[q, :n]))
        same = (np_(out["n_steps"])[q] == n and np_(out["status"])[q] == ref["status"][q]
                and np_(out["n_nodes"])[q] == ref["n_nodes"][q]
                and np.allclose(path[q, :n], rpath[q, :n], rtol=1e-12, atol=1e-13)
                and np.array_equal(np_(out["node_step"])[q, :ref["n_nodes"][q]],
                                   np.asarray(ref["node_step"])[q, :ref["n_nodes"][q]])
                and (ref.get("ctrl_idx") is None or np.array_equal(np_(out["ctrl_idx"])[q, :n], ref["ctrl_idx"][q, :n])))
        if not same:
            assert fl[q] != 0, f"{what}: extension {q} differs from the oracle without a report"
            diverged += 1
    assert diverged <= max_diverged, f"{what}: {diverged} reported extensions differ from the oracle"
    assert bitwise >= 0.98 * len(fl), f"{what}: only {bitwise} of {len(fl)} extensions bit-identical"
    return dict(clean=int((fl == 0).sum()), tie=int(((fl & 2) != 0).sum()), diverged=diverged)


@pytest.mark.parametrize("mname", ("test_env1", "test_env2"))
def test_steer_dwa_golden(mname):
    g = golden("steer_dwa")
    env, o = make_envs(fixture_map(mname))
    out = _extend(env, g[mname + "_from"], g[mname + "_target"])
    fl = np_(out["flags"])
    clean = fl == 0
    assert np.array_equal(np_(out["n_steps"])[clean], g[mname + "_nsteps"][clean])
    assert np.array_equal(np_(out["status"])[clean], g[mname + "_status"][clean])
    assert np.array_equal(np_(out["n_nodes"])[clean], g[mname + "_nnodes"][clean])
    assert clean.mean() >= 0.9
    for q in np.where(clean)[0]:
        n = g[mname + "_nsteps"][q]
        np.testing.assert_allclose(np_(out["path"])[q, :n], g[mname + "_path"][q, :n], rtol=1e-12, atol=1e-13)


def test_extend_vs_oracle_many():
    from crl_kino_b200.workloads import free_states, random_states
    env, o = make_envs(fixture_map("test_env1"))
    rng = np.random.default_rng(8)
    froms = free_states(lambda c: o.valid_state_batch(c)[1], rng, 300, 0.1, moving=True)
    targets = random_states(rng, 300)
    out = _extend(env, froms, targets)
    ref = orc.Dwa(o, dt=0.2, v_res=0.02).extend_batch(froms, targets, want_ctrl=True)
    _compare(out, ref, "test_env1", max_flagged=0.05, max_diverged=1)
    assert set(np.unique(ref["status"])) == {0, 1, 2}


def test_extend_dense_map_vs_oracle():
    """BASELINE config 4's shape at half its batch: 10,004 obstacles, ~4096-rollout windows, 512 extensions
    of the shared bench workload (workloads.cfg4_workload: nearest tree node -> accepted sample) against
    the threaded oracle: every unreported extension identical down to the chosen rollout of every control
    step and the last bit of every executed state."""
    from crl_kino_b200.workloads import DENSE_DWA, cfg4_workload
    Q = 512
    w = cfg4_workload(Q=Q, tree_nodes=1024, n_batches=1, seed=100)
    env, o = make_envs(w["rects"])
    tr, smp = w["trees"], w["batches"][0]
    idx, _ = orc.nearest_batch(np.ascontiguousarray(tr[:, :, 0]), np.ascontiguousarray(tr[:, :, 1]),
                               np.full(Q, tr.shape[1], dtype=np.int32), smp)
    froms = tr[np.arange(Q), idx]
    out = _extend(env, froms, smp, **DENSE_DWA)
    ref = orc.Dwa(o, **DENSE_DWA).extend_batch(froms, smp, want_ctrl=True)
    r = _compare(out, ref, "dense", max_flagged=0.15, max_diverged=2)
    assert r["clean"] >= 0.85 * Q
    assert int(ref["n_steps"].max()) >= 40 and set(np.unique(ref["status"])) >= {1, 2}


def test_extend_inactive_and_empty():
    import torch
    from crl_kino_b200 import ops
    from crl_kino_b200.policy.dwa import DWA
    env, o = make_envs(fixture_map("test_env1"))
    d = DWA(env)
    d.set_dwa(dt=0.2, v_res=0.02)
    s = np.array([[-5, -15, 0, 0, 0.0]] * 3)
    t = np.array([[10, 10, 0, 0, 0.0]] * 3)
    act = torch.tensor([1, 0, 1], dtype=torch.uint8, device="cuda")
    out = ops.extend_dwa(env.handle, d.params(), ops.extend_params(), s, t, active=act)
    ns = np_(out["n_steps"])
    assert ns[1] == 0 and ns[0] == ns[2] > 0 and np_(out["n_nodes"])[1] == 0
    out = ops.extend_dwa(env.handle, d.params(), ops.extend_params(), np.zeros((0, 5)), np.zeros((0, 5)))
    assert out["n_steps"].numel() == 0


@pytest.mark.parametrize("run,seed", ((0, 0), (2, 3)))
def test_rrt_class_reproduces_reference_run(run, seed):
    """The drop-in RRT with the reference's own RNG seeds rebuilds the reference's tree."""
    from crl_kino_b200.planner.rrt import RRT
    g = golden("plan_dwa")
    env, _ = make_envs(fixture_map(str(g[f"run{run}_map"])))
    p = RRT(env, max_iter=int(g[f"run{run}_max_iter"]))
    p.verbose = False
    p.set_start_and_goal(g[f"run{run}_start"], g[f"run{run}_goal"])
    random.seed(seed)
    np.random.seed(seed)
    path = p.planning()
    index = {id(n): i for i, n in enumerate(p.node_list)}
    parents = np.array([-1 if n.parent is None else index[id(n.parent)] for n in p.node_list])
    states = np.array([n.state for n in p.node_list])
    assert p.n_control_steps == int(g[f"run{run}_ctrl_steps"])
    assert np.array_equal(parents, g[f"run{run}_parents"])
    assert np.array_equal(np.array([len(n.path) for n in p.node_list]), g[f"run{run}_plen"])
    np.testing.assert_allclose(states, g[f"run{run}_node_states"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(np.array(path), g[f"run{run}_path"], rtol=1e-12, atol=1e-13)
    assert p.reach_exactly == bool(g[f"run{run}_reach"])


def test_rrt_api_errors():
    from crl_kino_b200.planner.rrt import RRT
    env, _ = make_envs(fixture_map("test_env1"))
    p = RRT(env)
    with pytest.raises(ValueError):
        p.planning()
    with pytest.raises(AssertionError):
        p.set_start_and_goal(np.array([0, -21.0, 0, 0, 0]), np.array([10, 10, 0, 0, 0.0]))


def test_pipelined_host_api_equals_synchronous():
    """BatchedExtender.submit_host / collect (two steps in flight) returns, step for step, the same
    arrays as step_host."""
    from crl_kino_b200 import ops
    from crl_kino_b200.planner.batched import BatchedExtender, Forest
    from crl_kino_b200.policy.dwa import DWA
    from crl_kino_b200.workloads import dense_map, free_states, random_states
    env, o = make_envs(dense_map(n_cells=2000, seed=3))
    rng = np.random.default_rng(8)
    Q, N = 96, 64
    nodes = free_states(lambda c: o.valid_state_batch(c)[1], rng, Q * N, 0.5)
    nodes[:, 3:] = 0
    forest = Forest(Q, N)
    forest.fill(nodes.reshape(Q, N, 5))
    d = DWA(env)
    d.set_dwa(dt=0.2, v_res=0.05, w_res=0.2)
    ext = BatchedExtender(env, forest, dwa=d)
    batches = [random_states(rng, Q, moving=True) for _ in range(5)]
    sync = [{k: v.copy() for k, v in ext.step_host(b).items()} for b in batches]
    got, pending = [], None
    for b in batches:
        t = ext.submit_host(b)
        if pending is not None:
            got.append({k: v.copy() for k, v in ext.collect(pending).items()})
        pending = t
    got.append({k: v.copy() for k, v in ext.collect(pending).items()})
    for a, b in zip(sync, got):
        ns = a["n_steps"]
        for k in ("n_steps", "status", "n_nodes", "nearest_idx", "flags"):
            assert np.array_equal(a[k], b[k]), k
        for q in range(Q):
            assert np.array_equal(a["path"][q, :ns[q]], b["path"][q, :ns[q]])
    # the device-resident form (submit / wait: inputs and results stay in HBM, pipe_depth steps in flight)
    tickets = [ext.submit(ops.as_dev(b)) for b in batches[:ext.pipe_depth]]
    for a, t in zip(sync, tickets):
        out = ext.wait(t)
        ns = a["n_steps"]
        assert np.array_equal(a["n_steps"], np_(out["n_steps"])) and np.array_equal(a["status"], np_(out["status"]))
        path = np_(out["path"])
        for q in range(Q):
            assert np.array_equal(a["path"][q, :ns[q]], path[q, :ns[q]])


@pytest.mark.parametrize("seed", range(10))
def test_rrt_class_contract_size(seed):
    """examples/test.py --case rrt as shipped (max_iter 500, rrt.py:35) through the drop-in RRT with
    the reference's own RNG seeding: the same tree as the unmodified reference (golden plan_dwa_full)."""
    from crl_kino_b200.planner.rrt import RRT
    g = golden("plan_dwa_full")
    env, _ = make_envs(fixture_map("test_env1"))
    p = RRT(env)
    p.verbose = False
    p.set_start_and_goal(g[f"run{seed}_start"], g[f"run{seed}_goal"])
    random.seed(seed)
    np.random.seed(seed)
    path = p.planning()
    index = {id(n): i for i, n in enumerate(p.node_list)}
    parents = np.array([-1 if n.parent is None else index[id(n.parent)] for n in p.node_list])
    states = np.array([n.state for n in p.node_list])
    same = len(states) == len(g[f"run{seed}_node_states"]) and np.array_equal(parents, g[f"run{seed}_parents"])
    if not same:
        assert p.flags_seen & 3, "tree differs from the reference without a near-tie / near-boundary report"
        pytest.skip("a reported near-tie / near-boundary decision went the other way")
    assert p.n_extend_calls == int(g[f"run{seed}_extensions"])
    assert p.n_control_steps == int(g[f"run{seed}_ctrl_steps"])
    np.testing.assert_allclose(states, g[f"run{seed}_node_states"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(np.array(path), g[f"run{seed}_path"], rtol=1e-12, atol=1e-13)
    assert p.reach_exactly == bool(g[f"run{seed}_reach"])
