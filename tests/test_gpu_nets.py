"""Actor MLP, estimator CNNs and RL steering vs goldens / oracle."""
import numpy as np
import pytest

from conftest import golden
from gpu_util import fixture_map, make_envs, np_
from oracle import oracle as orc
from test_oracle_golden import _seeded_actor

pytestmark = pytest.mark.gpu


def _policy(seed=0):
    from crl_kino_b200.env import DifferentialDriveGym
    from crl_kino_b200.policy.rl_policy import load_policy
    pol = load_policy(DifferentialDriveGym(), [1024, 512, 512, 512], None, seed=seed)
    return pol


def _estimators():
    from crl_kino_b200.estimator import load_estimator
    w = golden("estimator_weights")
    est = {k[4:]: w[k] for k in w.files if k.startswith("est.")}
    cls = {k[4:]: w[k] for k in w.files if k.startswith("cls.")}
    return load_estimator(est, cls), est, cls


def test_policy_forward_golden():
    from crl_kino_b200.policy.rl_policy import policy_forward
    g = golden("policy")
    pol = _policy(int(g["seed"]))
    Ws, bs = _seeded_actor(int(g["seed"]))
    assert all(np.array_equal(a, b) for a, b in zip(pol.actor.weights, Ws))
    a0 = policy_forward(pol, g["obs"], eps=0.0)
    np.testing.assert_allclose(a0, g["act_eps0"], rtol=1e-4, atol=1e-5)
    a1 = policy_forward(pol, g["obs"], eps=0.05, noise=g["noise"])
    np.testing.assert_allclose(a1, g["act_eps005"], rtol=1e-4, atol=1e-5)
    one = policy_forward(pol, g["obs"][7])
    assert one.shape == (1, 2)
    np.testing.assert_allclose(one[0], g["act_eps0"][7], rtol=1e-4, atol=1e-5)


def test_policy_forward_batches_vs_oracle():
    from crl_kino_b200.policy.rl_policy import policy_forward
    pol = _policy(3)
    rng = np.random.default_rng(0)
    for B in (1, 63, 64, 65, 1000):
        obs = np.concatenate([rng.normal(0, 5, (B, 4)), (rng.uniform(0, 1, (B, 729)) > 0.3).astype(float)], 1)
        noise = rng.normal(0, 1, (B, 2)).astype(np.float32)
        got = policy_forward(pol, obs, eps=0.05, noise=noise)
        ref = orc.actor_forward(pol.actor.weights, pol.actor.biases, obs, noise, 0.05)
        np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5)


def test_estimator_golden():
    from crl_kino_b200 import ops
    g = golden("estimator")
    (est, cls), esd, csd = _estimators()
    obs32 = ops.pad_obs(g["obs"])
    np.testing.assert_allclose(np_(est.forward_obs(obs32)), g["est"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(np_(cls.forward_obs(obs32)), g["cls"], rtol=1e-4, atol=1e-5)
    sp, ext = orc.estimator_inputs(g["obs"])
    np.testing.assert_allclose(np_(est(sp, ext))[:, 0], g["est"], rtol=1e-4, atol=1e-5)   # reference call form


def test_estimator_choose_parent_golden():
    from crl_kino_b200.planner.rrt_rl_estimator import RRT_RL_Estimator
    g = golden("estimator")
    (est, cls), _, _ = _estimators()
    env, _ = make_envs(fixture_map("test_env2"))
    p = RRT_RL_Estimator(env, _policy(0), est, cls)
    for k in range(3):
        nodes = [p.Node(s) for s in g[f"cp{k}_nodes"]]
        for t, i, c in zip(g[f"cp{k}_targets"], g[f"cp{k}_idx"], g[f"cp{k}_cost"]):
            gi, gc = p.choose_parent(nodes, p.Node(t))
            assert gi == i
            assert gc == pytest.approx(c, rel=1e-4)


@pytest.mark.parametrize("tag,eps", (("eps0", 0.0), ("eps005", 0.05)))
def test_steer_rl_golden(tag, eps):
    """RRT_RL.steer: closed loop with the float32 policy; states to 1e-4, flags exact
    unless the kernel reported a near-boundary decision."""
    from crl_kino_b200 import ops
    g = golden("steer_rl")
    env, _ = make_envs(fixture_map("test_env1"))
    pol = _policy(int(g["seed"]))
    out = ops.extend_rl(env.handle, pol.actor.handle, ops.extend_params(), g[tag + "_from"], g[tag + "_target"],
                        g[tag + "_noise"] if eps else None, eps, want_actions=True)
    fl = np_(out["flags"])
    ns = np_(out["n_steps"])
    for q in range(len(ns)):
        n_ref = int(g[tag + "_nsteps"][q])
        n = min(n_ref, ns[q])
        np.testing.assert_allclose(np_(out["actions"])[q, :n], g[tag + "_act"][q, :n], rtol=2e-3, atol=2e-4)
        np.testing.assert_allclose(np_(out["path"])[q, :n], g[tag + "_path"][q, :n], rtol=1e-4, atol=1e-4)
        if fl[q] == 0:
            assert ns[q] == n_ref
            assert np_(out["status"])[q] == g[tag + "_status"][q]
            assert np_(out["n_nodes"])[q] == g[tag + "_nnodes"][q]


def test_steer_rl_dense_map_golden():
    """RRT_RL.steer of the unmodified reference on BASELINE config 4's map (10,004 obstacles, golden
    steer_rl_dense): closed loop with the float32 policy, states to 1e-4, verdicts exact unless the
    kernel reported a near-boundary decision."""
    import os
    from crl_kino_b200 import ops
    from crl_kino_b200.env import DifferentialDriveEnv
    from crl_kino_b200.workloads import dense_map
    from conftest import GOLDEN
    if not os.path.exists(os.path.join(GOLDEN, "steer_rl_dense.npz")):
        pytest.skip("tests/golden/steer_rl_dense.npz not generated")
    g = golden("steer_rl_dense")
    env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi)
    env.set_obs(dense_map(n_cells=int(g["n_cells"])))
    pol = _policy(int(g["seed"]))
    out = ops.extend_rl(env.handle, pol.actor.handle, ops.extend_params(), g["from"], g["target"], None, 0.0,
                        want_actions=True)
    fl, ns = np_(out["flags"]), np_(out["n_steps"])
    for q in range(len(ns)):
        n_ref = int(g["nsteps"][q])
        n = min(n_ref, ns[q])
        np.testing.assert_allclose(np_(out["actions"])[q, :n], g["act"][q, :n], rtol=2e-3, atol=2e-4)
        np.testing.assert_allclose(np_(out["path"])[q, :n], g["path"][q, :n], rtol=1e-4, atol=1e-4)
        if fl[q] == 0:
            assert ns[q] == n_ref
            assert np_(out["status"])[q] == g["status"][q]
            assert np_(out["n_nodes"])[q] == g["nnodes"][q]


def test_extend_rl_vs_oracle_teacher_forced():
    """Per-step parity with teacher forcing: feed the kernel's own actions to the oracle dynamics."""
    from crl_kino_b200 import ops
    from crl_kino_b200.workloads import free_states, random_states
    env, o = make_envs(fixture_map("test_env2"))
    pol = _policy(1)
    rng = np.random.default_rng(2)
    froms = free_states(lambda c: o.valid_state_batch(c)[1], rng, 40, 0.5)
    targets = random_states(rng, 40)
    out = ops.extend_rl(env.handle, pol.actor.handle, ops.extend_params(), froms, targets, None, 0.0,
                        want_actions=True)
    ns, acts, path = np_(out["n_steps"]), np_(out["actions"]), np_(out["path"])
    for q in range(40):
        st = froms[q].copy()
        for k in range(ns[q]):
            obs = o.local_obs(st, targets[q])
            a = orc.actor_forward(pol.actor.weights, pol.actor.biases, obs)[0]
            np.testing.assert_allclose(acts[q, k], a, rtol=2e-3, atol=2e-4)
            st = o.motion_velocity(st, acts[q, k].astype(np.float64), 0.2)
            np.testing.assert_allclose(path[q, k], st, rtol=1e-9, atol=1e-11)
            code = 1 if not o.valid_state_check(st) else (2 if np.hypot(*(st[:2] - targets[q, :2])) <= 1.0 else 0)
            if k < ns[q] - 1:
                assert code == 0
        if np_(out["flags"])[q] == 0:
            assert code == np_(out["status"])[q] or (code == 0 and ns[q] == 50)


def test_extend_rl_large_batch_split_is_invisible():
    """Batches >= 2048 run as two halves on two streams over compacted row lists: every output must
    be bit-identical to running the same queries in small unsplit chunks (queries are independent),
    with exploration noise, an active mask and a noise cursor in play."""
    import torch
    from crl_kino_b200 import ops
    from crl_kino_b200.workloads import dense_map, free_states
    rects = dense_map(n_cells=3000, seed=6)
    env, o = make_envs(rects)
    pol = _policy(3)
    rng = np.random.default_rng(4)
    Q = 2304 + 37
    froms = free_states(lambda c: o.valid_state_batch(c)[1], rng, Q, 0.3, moving=True)
    targets = free_states(lambda c: o.valid_state_batch(c)[1], rng, Q, 0.3)
    noise = torch.as_tensor(rng.standard_normal((Q, 64, 2)).astype(np.float32)).cuda()
    active = torch.as_tensor((rng.uniform(size=Q) < 0.9).astype(np.uint8)).cuda()
    cur0 = torch.as_tensor(rng.integers(0, 10, Q).astype(np.int32)).cuda()
    xp = ops.extend_params()

    def run(lo, hi):
        cur = cur0[lo:hi].clone()
        out = ops.extend_rl(env.handle, pol.actor.handle, xp, froms[lo:hi], targets[lo:hi], noise[lo:hi].contiguous(),
                            0.05, active=active[lo:hi].contiguous(), want_actions=True, noise_cursor=cur)
        torch.cuda.synchronize()
        return {k: np_(v) for k, v in out.items() if k in ("path", "n_steps", "status", "node_step", "n_nodes",
                                                           "actions", "flags")}, np_(cur)
    whole, cur_w = run(0, Q)
    lo = 0
    for hi in list(range(700, Q, 700)) + [Q]:
        part, cur_p = run(lo, hi)
        ns = part["n_steps"]
        for k in ("n_steps", "status", "n_nodes", "flags"):
            assert np.array_equal(whole[k][lo:hi], part[k]), k
        assert np.array_equal(cur_w[lo:hi], cur_p)
        for q in range(hi - lo):                                  # rows past n_steps are not written
            assert np.array_equal(whole["path"][lo + q, :ns[q]], part["path"][q, :ns[q]])
            assert np.array_equal(whole["actions"][lo + q, :ns[q]], part["actions"][q, :ns[q]])
            assert np.array_equal(whole["node_step"][lo + q, :part["n_nodes"][q]], part["node_step"][q, :part["n_nodes"][q]])
        lo = hi
    assert (whole["n_steps"][np_(active) == 0] == 0).all() and whole["n_steps"].max() > 10


def test_policy_forward_large_batch_pair_kernel():
    """Batches >= 49,152 rows run the hidden layers on the CTA-pair GEMM (tcgen05.mma.cta_group::2,
    256 x 256 tiles, W halves shared between two SMs); smaller ones in the one-launch chain.  Both
    accumulate every output in the same order, so a large batch must equal, bit for bit, the same rows
    evaluated in small chunks -- and agree with the oracle's float64-checked forward pass."""
    import torch
    from crl_kino_b200 import ops
    pol = _policy(5)
    rng = np.random.default_rng(6)
    B = 49152 + 256
    obs = np.concatenate([rng.normal(0, 4, (B, 4)), (rng.uniform(0, 1, (B, 729)) > 0.4).astype(np.float64)], 1)
    o32 = ops.pad_obs(obs)
    noise = torch.as_tensor(rng.standard_normal((B, 2)).astype(np.float32)).cuda()
    big = np_(ops.policy_forward(pol.actor.handle, o32, noise, 0.05))
    for lo in (0, 4096, B - 1024):
        part = np_(ops.policy_forward(pol.actor.handle, o32[lo:lo + 1024].contiguous(), noise[lo:lo + 1024].contiguous(), 0.05))
        assert np.array_equal(big[lo:lo + 1024], part)
    ref = orc.actor_forward(pol.actor.weights, pol.actor.biases, obs[:256], np_(noise)[:256], 0.05)
    np.testing.assert_allclose(big[:256], ref, rtol=1e-4, atol=1e-5)


_CHAIN_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from crl_kino_b200 import ops
from crl_kino_b200.env import DifferentialDriveGym
from crl_kino_b200.policy.rl_policy import load_policy
pol = load_policy(DifferentialDriveGym(), [1024, 512, 512, 512], None, seed=5)
rng = np.random.default_rng(9)
B = 5000
obs = np.concatenate([rng.normal(0, 4, (B, 4)), (rng.uniform(0, 1, (B, 729)) > 0.4).astype(np.float64)], 1)
act = ops.policy_forward(pol.actor.handle, ops.pad_obs(obs)).cpu().numpy()
np.save(sys.argv[2], act)
"""


def test_mlp_chain_equals_one_launch_per_layer(tmp_path):
    """All hidden layers in one dataflow launch (k_mlp_chain: tiles of all layers pulled from one list,
    row-tile completion flags between layers) against one launch per layer (CRLK_MLP_CHAIN=0) and
    against the chain with 256-column tiles (CRLK_MLP_BN=256): the same products accumulated in the
    same k order, so the actions must agree bit for bit -- a missed dependency or a stale read between
    layers would not.  The switches are read once per process, hence the child processes."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for name, env in (("chain", {}), ("per_layer", {"CRLK_MLP_CHAIN": "0"}), ("chain_bn256", {"CRLK_MLP_BN": "256"})):
        f = str(tmp_path / (name + ".npy"))
        r = subprocess.run([sys.executable, "-c", _CHAIN_CHILD, root, f], env=dict(os.environ, **env),
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[name] = np.load(f)
    assert np.isfinite(outs["chain"]).all() and np.ptp(outs["chain"][:, 0]) > 0
    assert np.array_equal(outs["chain"], outs["per_layer"])
    assert np.array_equal(outs["chain"], outs["chain_bn256"])
