"""BatchedRRT (device-resident lock-step planner) vs the reference's full planning runs
and vs the oracle's planner port on synthetic queries."""
import numpy as np
import pytest

from conftest import golden
from gpu_util import fixture_map, make_envs, np_
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _planner_cls(kind):
    from crl_kino_b200.planner import batched
    return {"lockstep": batched.BatchedRRT, "fused": batched.FusedRRT}[kind]


@pytest.mark.parametrize("kind", ("lockstep", "fused"))
@pytest.mark.parametrize("runs", ((0, 1), (2,)))
def test_batched_rrt_reproduces_reference_runs(runs, kind):
    BatchedRRT = _planner_cls(kind)
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


@pytest.mark.parametrize("kind", ("lockstep", "fused"))
def test_batched_rrt_vs_oracle_dense_map(kind):
    BatchedRRT = _planner_cls(kind)
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


@pytest.mark.parametrize("mname", ("env1", "env2"))
def test_fused_rrt_benchmark_positions_vs_oracle(mname):
    """BASELINE config 3 (examples/benchmark.py:48-75): all 12 ordered (start, goal) pairs of the
    reference's benchmark_position_test_env{1,2}.pkl on test_env{1,2}.pkl, DWA steering with the
    RRT defaults, planned together in one launch; trees, parents, reach flags and final paths
    against the planner port (itself pinned to three full runs of the unmodified reference)."""
    from crl_kino_b200.planner.batched import FusedRRT
    from crl_kino_b200.workloads import sample_stream
    fx = golden("fixtures")
    env, o = make_envs(fx["map_test_" + mname])
    pos = fx["positions_" + mname]
    pairs = [(i, j) for i in range(4) for j in range(4) if i != j]       # benchmark.py:56-59
    starts = np.array([pos[i] for i, _ in pairs])
    goals = np.array([pos[j] for _, j in pairs])
    iters, S = 30, 400
    sb = env.get_bounds()["state_bounds"]
    streams = np.stack([sample_stream(sb, goals[q], S, seed=900 + q) for q in range(len(pairs))])
    p = FusedRRT(env, starts, goals, max_iter=iters)
    reached = p.plan(streams)
    flags = np_(p.flags)
    n_same = 0
    for q in range(len(pairs)):
        port = orc.PlannerPort(o, max_iter=iters)
        port.set_start_and_goal(starts[q], goals[q])
        port.planning(iter(streams[q]))
        ref_states = np.array([n.state for n in port.node_list])
        states, parents = p.tree(q)
        same = len(states) == len(ref_states) and np.allclose(states, ref_states, rtol=1e-7, atol=1e-9)
        if not same:
            assert flags[q] & 3, f"query {q} differs from the planner port without a near-tie/near-boundary report"
            continue
        index = {id(n): i for i, n in enumerate(port.node_list)}
        ref_par = np.array([-1 if n.parent is None else index[id(n.parent)] for n in port.node_list])
        assert np.array_equal(parents, ref_par)
        assert reached[q] == port.reach_exactly
        assert int(np_(p.cursor)[q]) == port.n_samples            # same number of draws consumed
        np.testing.assert_allclose(np.array(p.final_course(q)), np.array(port.path), rtol=1e-7, atol=1e-9)
        n_same += 1
    assert n_same >= len(pairs) - 1, f"only {n_same}/{len(pairs)} queries identical to the planner port"


def test_fused_rrt_equals_lockstep_cfg4_sizes():
    """One-launch planner vs the lock-step planner on the cfg4 map (10,004 obstacles, 64 x 64 DWA
    grid): identical trees, parents, path pools, iteration and sample counts for every query."""
    from crl_kino_b200.planner.batched import BatchedRRT, FusedRRT
    from crl_kino_b200.policy.dwa import DWA
    from crl_kino_b200.workloads import DENSE_DWA, dense_map, query_pairs, sample_stream
    rects = dense_map()
    env, o = make_envs(rects)
    Q, iters, S = 48, 8, 96
    s, gl = query_pairs(lambda c: o.valid_state_batch(c)[1], Q, seed=11)
    sb = env.get_bounds()["state_bounds"]
    streams = np.stack([sample_stream(sb, gl[q], S, seed=700 + q) for q in range(Q)])
    planners = []
    for cls in (BatchedRRT, FusedRRT):
        d = DWA(env)
        d.set_dwa(**DENSE_DWA)
        p = cls(env, s, gl, max_iter=iters, dwa=d)
        reached = p.plan(streams, check_every=4)
        planners.append((p, reached))
    (a, ra), (b, rb) = planners
    assert np.array_equal(ra, rb)
    assert np.array_equal(np_(a.forest.n_nodes), np_(b.forest.n_nodes))
    assert np.array_equal(np_(a.iters), np_(b.iters))
    assert int(np_(b.forest.n_nodes).sum()) > 2 * Q          # the runs did grow trees
    for q in range(Q):
        sa, pa = a.tree(q)
        sb_, pb = b.tree(q)
        assert np.array_equal(pa, pb)
        assert np.array_equal(sa, sb_)                       # same kernels, same operations: bit-identical
        assert np.array_equal(np.array(a.final_course(q)), np.array(b.final_course(q)))
    # a query that is not done when its stream ends reports cursor == S
    cur, it = np_(b.cursor), np_(b.iters)
    assert np.all((cur <= S) & (it <= iters))
    assert np.all(rb | (it == iters) | (cur == S))


def _policy(seed=0):
    from crl_kino_b200.env import DifferentialDriveGym
    from crl_kino_b200.policy.rl_policy import load_policy
    return load_policy(DifferentialDriveGym(), [1024, 512, 512, 512], None, seed=seed)


def _check_rl_tree(states, parents, path, g, kind):
    assert np.array_equal(parents, g[kind + "_parents"])
    np.testing.assert_allclose(states, g[kind + "_node_states"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(np.array(path), g[kind + "_path"], rtol=1e-4, atol=1e-4)


def test_batched_rrt_rl_reproduces_reference_run():
    """RRT_RL.planning (policy steering, eps = 0.05) with the recorded sample and noise streams."""
    from crl_kino_b200.planner.batched import BatchedRRT
    g = golden("plan_rl")
    env, _ = make_envs(fixture_map("test_env1"))
    start, goal = np.array([[-5, -15, 0, 0, 0.0]]), np.array([[10, 10, 0, 0, 0.0]])
    p = BatchedRRT(env, start, goal, max_iter=int(g["rl_max_iter"]), policy=_policy(int(g["seed"])), eps=0.05)
    noise = np.concatenate([g["rl_noise"], np.zeros((128, 2), dtype=np.float32)])[None]
    reached = p.plan(g["rl_samples"][None], check_every=4, noise=noise)
    states, parents = p.tree(0)
    _check_rl_tree(states, parents, p.final_course(0), g, "rl")
    assert int(p.noise_cursor[0].item()) == len(g["rl_noise"])          # one draw per executed step
    assert bool(reached[0]) == bool(g["rl_reach"])


def test_rrt_rl_estimator_class_reproduces_reference_run():
    """RRT_RL_Estimator (learned parent selection + policy steering), single-query drop-in class."""
    from crl_kino_b200.estimator import load_estimator
    from crl_kino_b200.planner.rrt_rl_estimator import RRT_RL_Estimator
    g = golden("plan_rl")
    w = golden("estimator_weights")
    est, cls = load_estimator({k[4:]: w[k] for k in w.files if k.startswith("est.")},
                              {k[4:]: w[k] for k in w.files if k.startswith("cls.")})
    env, _ = make_envs(fixture_map("test_env1"))
    p = RRT_RL_Estimator(env, _policy(int(g["seed"])), est, cls, max_iter=int(g["rl_est_max_iter"]))
    p.verbose = False
    p.set_start_and_goal(np.array([-5, -15, 0, 0, 0.0]), np.array([10, 10, 0, 0, 0.0]))
    it = iter(g["rl_est_samples"])
    p.sample = lambda goal: p.Node(next(it))
    p.set_noise_stream(np.concatenate([g["rl_est_noise"], np.zeros((64, 2), dtype=np.float32)]))
    path = p.planning()
    index = {id(n): i for i, n in enumerate(p.node_list)}
    parents = np.array([-1 if n.parent is None else index[id(n.parent)] for n in p.node_list])
    states = np.array([n.state for n in p.node_list])
    _check_rl_tree(states, parents, path, g, "rl_est")
    assert p.n_control_steps == len(g["rl_est_noise"])
    assert next(it, None) is None                                      # every recorded sample was consumed


@pytest.mark.parametrize("copies", (1, 3))
def test_batched_rrt_estimator_reproduces_reference_run(copies):
    """RRT_RL_Estimator.planning (learned parent selection for the sample AND for the retry, policy
    steering, eps = 0.05) for several queries in lock step: every copy of the recorded query rebuilds
    the reference's tree from the recorded sample and noise streams."""
    from crl_kino_b200.estimator import load_estimator
    from crl_kino_b200.planner.batched import BatchedRRT
    g = golden("plan_rl")
    w = golden("estimator_weights")
    est = load_estimator({k[4:]: w[k] for k in w.files if k.startswith("est.")},
                         {k[4:]: w[k] for k in w.files if k.startswith("cls.")})
    env, _ = make_envs(fixture_map("test_env1"))
    start = np.tile(np.array([[-5, -15, 0, 0, 0.0]]), (copies, 1))
    goal = np.tile(np.array([[10, 10, 0, 0, 0.0]]), (copies, 1))
    p = BatchedRRT(env, start, goal, max_iter=int(g["rl_est_max_iter"]), policy=_policy(int(g["seed"])), eps=0.05,
                   estimator=est)
    noise = np.concatenate([g["rl_est_noise"], np.zeros((128, 2), dtype=np.float32)])[None]
    samples = np.concatenate([g["rl_est_samples"], np.tile(goal[:1], (8, 1))])[None]
    reached = p.plan(np.tile(samples, (copies, 1, 1)), check_every=4, noise=np.tile(noise, (copies, 1, 1)))
    for q in range(copies):
        states, parents = p.tree(q)
        _check_rl_tree(states, parents, p.final_course(q), g, "rl_est")
        assert int(p.noise_cursor[q].item()) == len(g["rl_est_noise"])
        assert bool(reached[q]) == bool(g["rl_est_reach"])


def test_fused_rrt_device_sampling():
    """crlk_plan_dwa without sample streams draws RRT.sample's candidates itself (counter-based
    generator): the draws equal the host mirror bit for bit, and planning from the seed alone builds
    exactly the trees that planning from those materialised streams builds."""
    import torch
    from crl_kino_b200.planner.batched import FusedRRT
    from crl_kino_b200.workloads import device_sample_stream, query_pairs
    env, o = make_envs(fixture_map("test_env2"))
    Q, iters, S = 24, 12, 200
    s, gl = query_pairs(lambda c: o.valid_state_batch(c)[1], Q, seed=21)
    a = FusedRRT(env, s, gl, max_iter=iters)
    dev = a.sample_streams(S, seed=1234)
    host = device_sample_stream(env.get_bounds()["state_bounds"], gl, S, seed=1234)
    assert np.array_equal(np_(dev), host)
    frac_goal = np.mean(np.all(host == gl[:, None, :], axis=2))
    assert 0.03 < frac_goal < 0.09                                    # (5 + 1) / 101, rrt.py:147
    assert host[..., 0].min() >= -20 and host[..., 0].max() < 20 and abs(host[..., 2]).max() < np.pi
    ra = a.plan(None, seed=1234, max_samples=S)
    b = FusedRRT(env, s, gl, max_iter=iters)
    rb = b.plan(host)
    assert np.array_equal(ra, rb) and np.array_equal(np_(a.cursor), np_(b.cursor))
    assert torch.equal(a.forest.n_nodes, b.forest.n_nodes)
    for q in range(Q):
        sa, pa = a.tree(q)
        sb_, pb = b.tree(q)
        assert np.array_equal(sa, sb_) and np.array_equal(pa, pb)
    assert int(np_(a.forest.n_nodes).sum()) > 3 * Q


def _check_against_reference_runs(planner_cls, g, runs, env, **kw):
    """Plan the recorded reference runs `runs` of golden file `g` together (recorded sample streams)
    and compare trees, parents, reach flags and final paths with the reference's."""
    starts = np.array([g[f"run{r}_start"] for r in runs])
    goals = np.array([g[f"run{r}_goal"] for r in runs])
    S = max(len(g[f"run{r}_samples"]) for r in runs)
    streams = np.zeros((len(runs), S, 5))
    for i, r in enumerate(runs):
        s = g[f"run{r}_samples"]
        streams[i, :len(s)] = s
        streams[i, len(s):] = goals[i]          # never consumed: the query is done by then
    p = planner_cls(env, starts, goals, max_iter=int(g["max_iter"]), **kw)
    reached = p.plan(streams)
    flags = np_(p.flags) if hasattr(p, "flags") else np.zeros(len(runs), dtype=np.uint8)
    same = 0
    for i, r in enumerate(runs):
        states, parents = p.tree(i)
        ref = g[f"run{r}_node_states"]
        ok = len(states) == len(ref) and np.allclose(states, ref, rtol=1e-12, atol=1e-13)
        if not ok:
            assert flags[i] & 3, f"run {r} differs from the reference without a near-tie / near-boundary report"
            continue
        assert np.array_equal(parents, g[f"run{r}_parents"])
        assert reached[i] == bool(g[f"run{r}_reach"])
        np.testing.assert_allclose(np.array(p.final_course(i)), g[f"run{r}_path"], rtol=1e-12, atol=1e-13)
        if planner_cls.__name__ == "FusedRRT":
            assert int(np_(p.cursor)[i]) == len(g[f"run{r}_samples"])       # same number of draws consumed
        same += 1
    return same


def test_fused_rrt_reference_runs_contract_size():
    """BASELINE config 1 at the reference's size (examples/test.py:144-160; rrt.py:35 max_iter 500;
    seeds 0..9, up to 500 extensions / 5,458 control steps / 190 nodes per query): all ten reference
    runs planned in ONE launch of crlk_plan_dwa from the recorded sample streams."""
    from crl_kino_b200.planner.batched import FusedRRT
    g = golden("plan_dwa_full")
    env, _ = make_envs(fixture_map("test_env1"))
    same = _check_against_reference_runs(FusedRRT, g, list(range(10)), env)
    assert same >= 9, f"only {same}/10 runs identical to the reference"


@pytest.mark.parametrize("e", (1, 2))
def test_fused_rrt_benchmark_positions_vs_reference(e):
    """BASELINE config 3 (examples/benchmark.py:56-75): the 12 ordered (start, goal) pairs of
    benchmark_position_test_env{e}.pkl on test_env{e}.pkl, planned by the reference's own RRT
    (max_iter 500) -- here in one launch, same trees and paths."""
    from crl_kino_b200.planner.batched import FusedRRT
    g = golden("plan_cfg3")
    runs = [r for r in range(int(g["n_runs"])) if int(g[f"run{r}_env"]) == e]
    env, _ = make_envs(fixture_map(f"test_env{e}"))
    same = _check_against_reference_runs(FusedRRT, g, runs, env)
    assert same >= len(runs) - 1, f"only {same}/{len(runs)} runs identical to the reference"


def test_lockstep_rrt_reference_runs_contract_size():
    """The lock-step planner on three of the contract-size reference runs (seeds 1, 3, 6)."""
    from crl_kino_b200.planner.batched import BatchedRRT
    g = golden("plan_dwa_full")
    env, _ = make_envs(fixture_map("test_env1"))
    same = _check_against_reference_runs(BatchedRRT, g, [1, 3, 6], env)
    assert same == 3
