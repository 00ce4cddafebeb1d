"""BatchedRRT (device-resident lock-step planner) vs the reference's full planning runs
and vs the oracle's planner port on synthetic queries."""
import numpy as np
import pytest

from conftest import golden
from gpu_util import fixture_map, make_envs, np_
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("runs", ((0, 1), (2,)))
def test_batched_rrt_reproduces_reference_runs(runs):
    from crl_kino_b200.planner.batched import BatchedRRT
    g = golden("plan_dwa")
    env, _ = make_envs(fixture_map(str(g[f"run{runs[0]}_map"])))
    starts = np.array([g[f"run{r}_start"] for r in runs])
    goals = np.array([g[f"run{r}_goal"] for r in runs])
    S = max(len(g[f"run{r}_samples"]) for r in runs)
    streams = np.zeros((len(runs), S, 5))
    for i, r in enumerate(runs):
        s = g[f"run{r}_samples"]
        streams[i, :len(s)] = s
        streams[i, len(s):] = goals[i]          # never consumed: the query is done by then
    p = BatchedRRT(env, starts, goals, max_iter=int(g[f"run{runs[0]}_max_iter"]))
    reached = p.plan(streams, check_every=4)
    for i, r in enumerate(runs):
        states, parents = p.tree(i)
        assert np.array_equal(parents, g[f"run{r}_parents"])
        np.testing.assert_allclose(states, g[f"run{r}_node_states"], rtol=1e-9, atol=1e-11)
        assert reached[i] == bool(g[f"run{r}_reach"])
        path = np.array(p.final_course(i))
        np.testing.assert_allclose(path, g[f"run{r}_path"], rtol=1e-9, atol=1e-11)


def test_batched_rrt_vs_oracle_dense_map():
    from crl_kino_b200.planner.batched import BatchedRRT
    from crl_kino_b200.policy.dwa import DWA
    from crl_kino_b200.workloads import dense_map, query_pairs, sample_stream
    rects = dense_map(n_cells=3000, seed=2)
    env, o = make_envs(rects)
    Q, iters, S = 12, 6, 160
    s, gl = query_pairs(lambda c: o.valid_state_batch(c)[1], Q, seed=5)
    sb = env.get_bounds()["state_bounds"]
    streams = np.stack([sample_stream(sb, gl[q], S, seed=50 + q) for q in range(Q)])
    kw = dict(dt=0.2, v_res=0.05, w_res=0.2)
    d = DWA(env)
    d.set_dwa(**kw)
    p = BatchedRRT(env, s, gl, max_iter=iters, dwa=d)
    reached = p.plan(streams, check_every=2)
    n_ok = 0
    for q in range(Q):
        port = orc.PlannerPort(o, max_iter=iters, dwa_kwargs=kw)
        port.set_start_and_goal(s[q], gl[q])
        port.planning(iter(streams[q]))
        ref_states = np.array([n.state for n in port.node_list])
        states, parents = p.tree(q)
        if len(states) != len(ref_states) or not np.allclose(states, ref_states, rtol=1e-7, atol=1e-9):
            continue        # near-tie / near-boundary divergence (bounded below)
        index = {id(n): i for i, n in enumerate(port.node_list)}
        ref_par = np.array([-1 if n.parent is None else index[id(n.parent)] for n in port.node_list])
        assert np.array_equal(parents, ref_par)
        assert reached[q] == port.reach_exactly
        np.testing.assert_allclose(np.array(p.final_course(q)), np.array(port.path), rtol=1e-7, atol=1e-9)
        n_ok += 1
    assert n_ok >= Q - 1, f"only {n_ok}/{Q} queries identical to the planner port"
