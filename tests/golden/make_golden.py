#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by importing the UNMODIFIED
reference (read-only at /root/reference) with the tiny import stubs in
tests/golden/_stubs (matplotlib, seaborn, gym, transforms3d are not installed).

Runs only in the authoring container (the reference does not travel to the GPU
box); its outputs (*.npz) are committed.  Usage:

    python tests/golden/make_golden.py [--only NAME ...]

Nothing here is read at test time except the .npz files it wrote.
"""
import argparse
import contextlib
import io
import os
import pickle
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CRLK_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "_stubs"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from crl_kino.env import DifferentialDriveEnv, DifferentialDriveGym  # noqa: E402
from crl_kino.planner.rrt import RRT  # noqa: E402
from crl_kino.planner.rrt_rl import RRT_RL  # noqa: E402
from crl_kino.planner.rrt_rl_estimator import RRT_RL_Estimator  # noqa: E402
from crl_kino.policy.dwa import DWA  # noqa: E402
from crl_kino.policy.net import Actor, Critic  # noqa: E402
from crl_kino.policy.rl_policy import policy_forward  # noqa: E402
from crl_kino.estimator.load_estimator import load_estimator  # noqa: E402
from crl_kino.utils.rigid import RectObs, RectRobot  # noqa: E402
import crl_kino.policy.dwa as dwa_mod  # noqa: E402
from tianshou.policy import DDPGPolicy  # noqa: E402


DATA = os.path.join(os.path.dirname(os.path.dirname(HERE)), "data")   # inputs shared with bench.py / examples


def save(name, **arrs):
    path = os.path.join(DATA if name in ("fixtures", "estimator_weights") else HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"wrote {path}: {os.path.getsize(path)/1024:.1f} KiB")


def load_maps():
    d = os.path.join(REF, "data", "obstacles")
    maps = {}
    for n in ("test_env1", "test_env2", "training_env"):
        maps[n] = pickle.load(open(os.path.join(d, n + ".pkl"), "rb"))
    for i, m in enumerate(pickle.load(open(os.path.join(d, "obs_list_list.pkl"), "rb"))):
        maps[f"obs_list_{i}"] = m
    return maps


def rects_of(obs_list):
    return np.array([o.rect.reshape(4) for o in obs_list], dtype=np.float32)


def make_env(obs_list):
    env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi)
    env.set_obs(obs_list)
    return env


def random_states(rng, n, speed=True):
    s = np.zeros((n, 5))
    s[:, 0] = rng.uniform(-20, 20, n)
    s[:, 1] = rng.uniform(-20, 20, n)
    s[:, 2] = rng.uniform(-np.pi, np.pi, n)
    if speed:
        s[:, 3] = rng.uniform(-0.1, 1.0, n)
        s[:, 4] = rng.uniform(-np.pi, np.pi, n)
    return s


# ---------------------------------------------------------------- fixtures

def g_fixtures():
    """Input fixtures defined by the reference's data files, as flat arrays."""
    maps = load_maps()
    out = {"map_" + k: rects_of(v) for k, v in maps.items()}
    ex = os.path.join(REF, "examples")
    out["positions_env1"] = pickle.load(open(os.path.join(ex, "benchmark_position_test_env1.pkl"), "rb"))
    out["positions_env2"] = pickle.load(open(os.path.join(ex, "benchmark_position_test_env2.pkl"), "rb"))
    gym = DifferentialDriveGym(make_env(maps["test_env1"]))
    out["sample_positions"] = gym.sample_positions
    out["action_low"] = gym.action_space.low
    out["action_high"] = gym.action_space.high
    out["robot_rect"] = gym.robot_env.rigid_robot.rect.reshape(4)
    out["state_bounds"] = gym.robot_env.get_bounds()["state_bounds"]
    save("fixtures", **out)
    nd = os.path.join(REF, "data", "net", "estimator")
    w = {}
    for tag, f in (("est", "dist_est_statedict.pth"), ("cls", "classifier_statedict.pth")):
        sd = torch.load(os.path.join(nd, f), map_location="cpu")
        for k, v in sd.items():
            w[f"{tag}.{k}"] = v.numpy()
    save("estimator_weights", **w)


# ------------------------------------------------------------ numerics probes

def g_numerics():
    """np.arange / np.sum / normalize_angle behaviour the restatement mirrors."""
    from crl_kino.env.differential_env import normalize_angle
    rng = np.random.default_rng(11)
    starts = rng.uniform(-3.2, 1.0, 64)
    spans = rng.uniform(0.0, 1.3, 64)
    steps = rng.choice([0.02, 0.01, 0.1, 0.4 / 63, 0.4 * np.pi / 63, 0.05], 64)
    lens, a5 = [], []
    for s, sp, st in zip(starts, spans, steps):
        a = np.arange(s, s + sp + st, st)
        lens.append(len(a))
        a5.append(np.concatenate([a[:3], a[-2:]]) if len(a) >= 5 else np.zeros(5))
    sums_n = np.array([1, 5, 7, 8, 9, 63, 64, 99, 128, 129, 224, 255, 256, 294, 1000, 4096, 4225, 8191, 8200])
    sums, sums_in = [], []
    for n in sums_n:
        tab = rng.uniform(0, 3.2, (n, 3))
        sums_in.append(tab[:, 1].copy())
        sums.append(np.sum(tab[:, 1]))
    ang = np.concatenate([rng.uniform(-20, 20, 200), [0.0, np.pi, -np.pi, 2 * np.pi, -2 * np.pi, 3 * np.pi]])
    save("numerics", ar_start=starts, ar_span=spans, ar_step=steps, ar_len=np.array(lens), ar_samp=np.array(a5),
         sum_n=sums_n, sum_val=np.array(sums), sum_in=np.concatenate(sums_in),
         ang_in=ang, ang_out=np.array([normalize_angle(np.float64(a)) for a in ang]))


# ------------------------------------------------------------------ dynamics

def g_motion():
    rng = np.random.default_rng(1)
    env = make_env([])
    n = 200
    s = random_states(rng, n)
    cmd = np.stack([rng.uniform(-0.3, 1.2, n), rng.uniform(-4, 4, n)], 1)
    mv = np.array([env.motion_velocity(s[i], cmd[i], 0.2) for i in range(n)])
    dwa = DWA(env)
    dwa.set_dwa(dt=0.2, v_res=0.02)
    traj = np.array([dwa.calc_trajectory(s[i], cmd[i]) for i in range(n)])
    dwa2 = DWA(env)  # defaults: dt 0.1 -> 12 rows
    traj2 = np.array([dwa2.calc_trajectory(s[i], cmd[i]) for i in range(40)])
    acc = np.stack([rng.uniform(-1, 1, n), rng.uniform(-3, 3, n)], 1)
    mo = np.array([env.motion(s[i], acc[i], 0.5) for i in range(n)])
    dw = np.array([env.dynamic_window(s[i], 0.2) for i in range(n)])
    save("motion", states=s, cmd=cmd, mv=mv, traj=traj, traj_dt01=traj2, acc=acc, motion=mo, dw=dw)


# ----------------------------------------------------------------------- DWA

class ArgminTap:
    """Records the argument of the np.argmin call inside DWA.control."""

    def __init__(self):
        self.last = None
        self._orig = np.argmin

    def __enter__(self):
        def tap(a, *k, **kw):
            self.last = np.array(a)
            return self._orig(a, *k, **kw)
        dwa_mod.np.argmin = tap
        return self

    def __exit__(self, *exc):
        dwa_mod.np.argmin = self._orig


def g_dwa():
    rng = np.random.default_rng(2)
    env = make_env(load_maps()["test_env1"])
    cases = []
    # (name, dwa kwargs, n)
    confs = [("rrt", dict(dt=0.2, v_res=0.02), 60),
             ("default", dict(), 6),
             ("dense", dict(dt=0.2, v_res=0.4 / 63, w_res=0.4 * np.pi / 63), 4)]
    out = {}
    for name, kw, n in confs:
        dwa = DWA(env)
        dwa.set_dwa(**kw)
        s = random_states(rng, n)
        s[: n // 4, 3:] = 0.0                       # at rest (tree roots)
        s[n // 4: n // 2, 3] = 1.0                  # at the v limit
        g = random_states(rng, n, speed=False)
        optv, idx, R, cmin, trajs, csums = [], [], [], [], [], []
        for i in range(n):
            with ArgminTap() as tap:
                v, traj = dwa.control(s[i], g[i])
            cs = tap.last
            optv.append(v); idx.append(int(np.argmin(cs))); R.append(len(cs)); cmin.append(cs.min())
            trajs.append(traj)
            csums.append(cs)
            srt = np.sort(cs)
            cases.append(srt[1] - srt[0])
        out[f"{name}_states"] = s
        out[f"{name}_goals"] = g
        out[f"{name}_optv"] = np.array(optv)
        out[f"{name}_idx"] = np.array(idx)
        out[f"{name}_R"] = np.array(R)
        out[f"{name}_cmin"] = np.array(cmin)
        out[f"{name}_traj"] = np.array(trajs)
        out[f"{name}_csum"] = np.concatenate(csums[:8])   # full cost vectors of the first 8
    print("dwa: smallest best-vs-second gap", min(cases))
    save("dwa", **out)


def g_dwa_clearance():
    """DWA with the clearance term enabled = reference dwa.py with lines 65-68
    un-commented (a patched COPY executed from memory; the file on disk is not
    touched and not stored)."""
    src = open(os.path.join(REF, "crl_kino", "policy", "dwa.py")).read()
    a = "                # for i in range(0,len(traj),self.skip):\n"
    assert a in src
    src = src.replace("                # for i in range(0,len(traj),self.skip):", "                for i in range(0,len(traj),self.skip):")
    src = src.replace("                #     x = traj[i]", "                    x = traj[i]")
    src = src.replace("                #     dis = max(self.robot_env.get_clearance(x), 0.001)", "                    dis = max(self.robot_env.get_clearance(x), 0.001)")
    src = src.replace("                #     min_dis = min(dis, min_dis)", "                    min_dis = min(dis, min_dis)")
    ns = {}
    exec(compile(src, "dwa_patched", "exec"), ns)
    DWAc = ns["DWA"]
    rng = np.random.default_rng(3)
    maps = load_maps()
    out = {}
    for mname in ("test_env1", "test_env2"):
        env = make_env(maps[mname])
        dwa = DWAc(env)
        dwa.set_dwa(dt=0.2, v_res=0.1, w_res=0.4)     # small grid: the clearance term is slow
        n = 10
        s = []
        while len(s) < n:
            c = random_states(rng, 1)[0]
            if env.get_clearance(c) > 0.2 and env.get_clearance(c) < 3.0:
                s.append(c)
        s = np.array(s)
        g = random_states(rng, n, speed=False)
        optv, idx, R, cmin = [], [], [], []
        orig = np.argmin
        for i in range(n):
            rec = {}

            def tap(a, *k, **kw):
                rec["a"] = np.array(a)
                return orig(a, *k, **kw)
            ns["np"].argmin = tap
            try:
                v, traj = dwa.control(s[i], g[i])
            finally:
                ns["np"].argmin = orig
            cs = rec["a"]
            optv.append(v); idx.append(int(orig(cs))); R.append(len(cs)); cmin.append(cs.min())
        out[f"{mname}_states"] = s
        out[f"{mname}_goals"] = g
        out[f"{mname}_optv"] = np.array(optv)
        out[f"{mname}_idx"] = np.array(idx)
        out[f"{mname}_R"] = np.array(R)
        out[f"{mname}_cmin"] = np.array(cmin)
    save("dwa_clearance", **out)


# ------------------------------------------------------------------ geometry

def g_geometry():
    rng = np.random.default_rng(4)
    maps = load_maps()
    out = {}
    # pairwise dis2obs near each obstacle (so that many cases are close / colliding)
    robot = RectRobot(np.array([[-0.7, -0.5], [0.7, 0.5]]))
    poses, rects, dis = [], [], []
    for mname in ("test_env1", "test_env2", "training_env", "obs_list_0"):
        for o in maps[mname]:
            l, b, r, t = o.rect.reshape(4)
            for _ in range(24):
                p = np.array([rng.uniform(l - 2.0, r + 2.0), rng.uniform(b - 2.0, t + 2.0),
                              rng.uniform(-np.pi, np.pi)])
                robot.pose = p
                poses.append(p); rects.append(o.rect.reshape(4)); dis.append(robot.dis2obs(o))
    # small obstacles (incl. strictly inside the footprint: NOT a collision)
    for _ in range(300):
        c = rng.uniform(-1.5, 1.5, 2)
        wh = rng.uniform(0.05, 0.4, 2)
        o = RectObs(np.array([c - wh / 2, c + wh / 2]))
        p = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(-np.pi, np.pi)])
        robot.pose = p
        poses.append(p); rects.append(o.rect.reshape(4)); dis.append(robot.dis2obs(o))
    # axis-aligned poses (zero direction components -> the 1e300 branch)
    for th in (0.0, np.pi / 2, np.pi, -np.pi / 2):
        for _ in range(40):
            o = maps["test_env1"][rng.integers(4, 15)]
            l, b, r, t = o.rect.reshape(4)
            p = np.array([rng.uniform(l - 1.5, r + 1.5), rng.uniform(b - 1.5, t + 1.5), th])
            robot.pose = p
            poses.append(p); rects.append(o.rect.reshape(4)); dis.append(robot.dis2obs(o))
    out["pair_pose"] = np.array(poses)
    out["pair_rect"] = np.array(rects, dtype=np.float32)
    out["pair_dis"] = np.array(dis, dtype=np.float64)
    print("pairs:", len(dis), "colliding:", int(np.sum(np.array(dis) == -1)))
    # whole-map validity / clearance
    for mname in ("test_env1", "test_env2", "training_env", "obs_list_3"):
        env = make_env(maps[mname])
        s = random_states(rng, 150)
        out[f"{mname}_states"] = s
        out[f"{mname}_valid"] = np.array([env.valid_state_check(x) for x in s])
        out[f"{mname}_clear"] = np.array([env.get_clearance(x) for x in s], dtype=np.float64)
        pts = rng.uniform(-22, 22, (400, 2))
        out[f"{mname}_pts"] = pts
        out[f"{mname}_pts_valid"] = env.valid_point_check(pts)
    save("geometry", **out)


def g_obs():
    rng = np.random.default_rng(5)
    maps = load_maps()
    out = {}
    for mname in ("test_env1", "test_env2"):
        env = make_env(maps[mname])
        gym = DifferentialDriveGym(env)
        s = random_states(rng, 40)
        s[:4, 2:] = 0.0
        s[0, :2] = [-5, -15]
        s[1, :2] = [10, 10]
        g = random_states(rng, 40, speed=False)
        o = []
        for i in range(40):
            gym.state = s[i]; gym.goal = g[i]
            o.append(gym._obs())
        out[f"{mname}_states"] = s
        out[f"{mname}_goals"] = g
        out[f"{mname}_obs"] = np.array(o)
    save("obs", **out)


# ------------------------------------------------------------------- planner

def g_nearest():
    rng = np.random.default_rng(6)
    out = {}
    for n in (1, 7, 100, 1000, 5000):
        nodes = [RRT.Node(x) for x in random_states(rng, n)]
        tg = random_states(rng, 16)
        idx, dist = [], []
        for t in tg:
            i, d = RRT.choose_parent(nodes, RRT.Node(t))
            idx.append(i); dist.append(d)
        out[f"n{n}_xy"] = np.array([x.state[:2] for x in nodes])
        out[f"n{n}_targets"] = tg[:, :2]
        out[f"n{n}_idx"] = np.array(idx)
        out[f"n{n}_dist"] = np.array(dist)
    # exact ties: duplicated nodes -> first index wins
    base = random_states(rng, 50)
    nodes = [RRT.Node(x) for x in np.concatenate([base, base])]
    tg = random_states(rng, 8)
    out["tie_xy"] = np.array([x.state[:2] for x in nodes])
    out["tie_targets"] = tg[:, :2]
    out["tie_idx"] = np.array([RRT.choose_parent(nodes, RRT.Node(t))[0] for t in tg])
    save("nearest", **out)


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def pack_steer(planner, from_state, target):
    planner.node_list = []
    frm = planner.Node(from_state)
    new = planner.steer(frm, planner.Node(target))
    # recover: executed states = concatenated segment paths (+ the unfinished one is lost), so re-run manually
    return new


def g_steer_dwa():
    """RRT.steer (rrt.py:99-135) from random valid states; records per-step
    states, the emitted nodes and the termination cause."""
    rng = np.random.default_rng(7)
    maps = load_maps()
    out = {}
    for mname, n in (("test_env1", 14), ("test_env2", 10)):
        env = make_env(maps[mname])
        planner = RRT(env)
        froms, targets, paths, nsteps, status, nnodes, node_states = [], [], [], [], [], [], []
        while len(froms) < n:
            f = random_states(rng, 1)[0]
            f[3:] = rng.choice([0.0, 1.0]) * f[3:]
            if not env.valid_state_check(f):
                continue
            t = random_states(rng, 1)[0]
            d = np.linalg.norm(f[:2] - t[:2])
            if not (1.0 <= d < 20):
                continue
            # instrument: record every executed state via motion_velocity calls at steer level
            rec = []
            orig_valid = env.valid_state_check

            def tap(state, _rec=rec, _o=orig_valid):
                r = _o(state)
                _rec.append((np.array(state), r))
                return r
            env.valid_state_check = tap
            planner.node_list = []
            try:
                new = planner.steer(planner.Node(f), planner.Node(t))
            finally:
                env.valid_state_check = orig_valid
            p = np.zeros((50, 5))
            for k, (st, _) in enumerate(rec):
                p[k] = st
            st_code = 0
            if not rec[-1][1]:
                st_code = 1
            elif np.linalg.norm(rec[-1][0][:2] - t[:2]) <= 1.0:
                st_code = 2
            ns = np.zeros((3, 5))
            for k, nd in enumerate(new):
                ns[k] = nd.state
            froms.append(f); targets.append(t); paths.append(p); nsteps.append(len(rec))
            status.append(st_code); nnodes.append(len(new)); node_states.append(ns)
        out[f"{mname}_from"] = np.array(froms)
        out[f"{mname}_target"] = np.array(targets)
        out[f"{mname}_path"] = np.array(paths)
        out[f"{mname}_nsteps"] = np.array(nsteps)
        out[f"{mname}_status"] = np.array(status)
        out[f"{mname}_nnodes"] = np.array(nnodes)
        out[f"{mname}_node_states"] = np.array(node_states)
        print(mname, "steer status", status, "steps", nsteps)
    save("steer_dwa", **out)


def tap_samples(planner):
    """Wrap planner.sample so every draw is recorded."""
    rec = []
    orig = planner.sample

    def tapped(goal):
        n = orig(goal)
        rec.append(n.state.copy())
        return n
    planner.sample = tapped
    return rec


def pack_tree(planner):
    nodes = planner.node_list
    index = {id(n): i for i, n in enumerate(nodes)}
    states = np.array([n.state for n in nodes])
    parents = np.array([-1 if n.parent is None else index[id(n.parent)] for n in nodes])
    plen = np.array([len(n.path) for n in nodes])
    return states, parents, plen


def g_plan_dwa():
    """Full RRT.planning runs (rrt.py:51-89) with recorded sample streams."""
    maps = load_maps()
    out = {}
    runs = [("test_env1", [-5, -15, 0, 0, 0], [10, 10, 0, 0, 0], 0, 40),
            ("test_env1", [-5, -15, 0, 0, 0], [10, 10, 0, 0, 0], 1, 40),
            ("test_env2", [-15.0, 17, 0, 0, 0], [-7.6, -7.3, 0, 0, 0], 3, 40)]
    for k, (mname, start, goal, seed, iters) in enumerate(runs):
        env = make_env(maps[mname])
        planner = RRT(env, max_iter=iters)
        planner.set_start_and_goal(np.array(start, dtype=float), np.array(goal, dtype=float))
        rec = tap_samples(planner)
        random.seed(seed); np.random.seed(seed)
        steps = {"n": 0}
        oc = planner.dwa.control

        def ctrl(x, g, _s=steps, _oc=oc):
            _s["n"] += 1
            return _oc(x, g)
        planner.dwa.control = ctrl
        tic = time.time()
        with quiet():
            path = planner.planning()
        el = time.time() - tic
        states, parents, plen = pack_tree(planner)
        out[f"run{k}_map"] = np.array(mname)
        out[f"run{k}_start"] = np.array(start, dtype=float)
        out[f"run{k}_goal"] = np.array(goal, dtype=float)
        out[f"run{k}_max_iter"] = np.array(iters)
        out[f"run{k}_samples"] = np.array(rec)
        out[f"run{k}_node_states"] = states
        out[f"run{k}_parents"] = parents
        out[f"run{k}_plen"] = plen
        out[f"run{k}_path"] = np.array(path)
        out[f"run{k}_reach"] = np.array(planner.reach_exactly)
        out[f"run{k}_ctrl_steps"] = np.array(steps["n"])
        out[f"run{k}_ref_seconds"] = np.array(el)
        print(f"plan_dwa run{k}: {len(states)} nodes, {len(rec)} samples, {steps['n']} control steps, "
              f"reach={planner.reach_exactly}, {el:.1f}s")
    save("plan_dwa", **out)


# ---------------------------------------------------------------- RL pieces

LAYER = [1024, 512, 512, 512]


def make_policy(seed):
    """Seeded random-init actor of the shipped architecture (the trained
    policy.pth is missing from the reference checkout)."""
    gym = DifferentialDriveGym()
    torch.manual_seed(seed)
    rngk = [gym.action_space.low, gym.action_space.high]
    actor = Actor(LAYER, (733,), (2,), rngk, "cpu")
    critic = Critic(LAYER, (733,), (2,), "cpu")
    pol = DDPGPolicy(actor, None, critic, None, action_range=rngk)
    return pol, actor


def actor_arrays(actor):
    Ws = [m.weight.detach().numpy() for m in actor.model if isinstance(m, torch.nn.Linear)]
    bs = [m.bias.detach().numpy() for m in actor.model if isinstance(m, torch.nn.Linear)]
    return Ws, bs


def g_policy():
    rng = np.random.default_rng(8)
    maps = load_maps()
    env = make_env(maps["test_env1"])
    gym = DifferentialDriveGym(env)
    pol, actor = make_policy(0)
    Ws, bs = actor_arrays(actor)
    s = random_states(rng, 64)
    g = random_states(rng, 64, speed=False)
    obs = []
    for i in range(64):
        gym.state = s[i]; gym.goal = g[i]
        obs.append(gym._obs())
    obs = np.array(obs)
    act0 = policy_forward(pol, obs, eps=0.0)
    torch.manual_seed(123)
    act_eps = policy_forward(pol, obs, eps=0.05)
    torch.manual_seed(123)
    noise = torch.randn(size=(64, 2)).numpy()
    # weight fingerprint so that a test can confirm the seeded re-creation matches
    fp = np.array([float(np.sum(W.astype(np.float64))) for W in Ws] + [float(np.sum(b.astype(np.float64))) for b in bs])
    save("policy", obs=obs, act_eps0=act0, act_eps005=act_eps, noise=noise, weight_fingerprint=fp,
         seed=np.array(0))


def g_estimator():
    rng = np.random.default_rng(9)
    maps = load_maps()
    nd = os.path.join(REF, "data", "net", "estimator")
    est, cls = load_estimator(os.path.join(nd, "dist_est_statedict.pth"),
                              os.path.join(nd, "classifier_statedict.pth"))
    env = make_env(maps["test_env2"])
    gym = DifferentialDriveGym(env)
    s = random_states(rng, 48)
    g = random_states(rng, 48, speed=False)
    obs = []
    for i in range(48):
        gym.state = s[i]; gym.goal = g[i]
        obs.append(gym._obs())
    inputs = torch.tensor(np.array(obs), dtype=torch.float32)
    inputs[:, 1] = torch.norm(inputs[:, :2], dim=1)
    spatial = inputs[:, 4:].view([-1, 1, 27, 27])
    ext = inputs[:, 1:4]
    with torch.no_grad():
        e = est(spatial, ext)[:, 0].numpy()
        c = cls(spatial, ext)[:, 0].numpy()
    # choose_parent of RRT_RL_Estimator on random trees
    pol, _ = make_policy(0)
    planner = RRT_RL_Estimator(env, pol, est, cls)
    out = {}
    for k, n in enumerate((5, 40, 120)):
        nodes = [planner.Node(x) for x in random_states(rng, n)]
        tg = random_states(rng, 6)
        res = [planner.choose_parent(nodes, planner.Node(t)) for t in tg]
        out[f"cp{k}_nodes"] = np.array([x.state for x in nodes])
        out[f"cp{k}_targets"] = tg
        out[f"cp{k}_idx"] = np.array([r[0] for r in res])
        out[f"cp{k}_cost"] = np.array([float(r[1]) for r in res])
    save("estimator", obs=np.array(obs), est=e, cls=c, **out)


def g_steer_rl():
    """RRT_RL.steer (rrt_rl.py:22-63), seeded random-init policy, eps = 0 and
    eps = 0.05 with a recorded noise stream."""
    rng = np.random.default_rng(10)
    maps = load_maps()
    env = make_env(maps["test_env1"])
    pol, actor = make_policy(0)
    planner = RRT_RL(env, pol)
    import crl_kino.planner.rrt_rl as rl_mod
    out = {}
    for tag, eps in (("eps0", 0.0), ("eps005", 0.05)):
        froms, targets, paths, nsteps, status, nnodes, noises, acts = [], [], [], [], [], [], [], []
        while len(froms) < 8:
            f = random_states(rng, 1)[0]
            f[3:] = 0
            if env.get_clearance(f) < 1.0:
                continue
            t = random_states(rng, 1)[0]
            if not (3.0 <= np.linalg.norm(f[:2] - t[:2]) < 20):
                continue
            rec, nrec, arec = [], [], []
            orig_pf = rl_mod.policy_forward
            orig_randn = torch.randn

            def randn_tap(*a, **k):
                r = orig_randn(*a, **k)
                nrec.append(r.numpy().copy())
                return r

            def pf(policy, obs, info=None, eps=0.0, _eps=eps):
                torch.randn = randn_tap
                try:
                    a = orig_pf(policy, obs, info=info, eps=_eps)
                finally:
                    torch.randn = orig_randn
                arec.append(a[0].copy())
                return a
            rl_mod.policy_forward = pf
            orig_valid = env.valid_state_check

            def tap(state, _rec=rec, _o=orig_valid):
                r = _o(state)
                _rec.append((np.array(state), r))
                return r
            env.valid_state_check = tap
            planner.node_list = []
            try:
                new = planner.steer(planner.Node(f), planner.Node(t))
            finally:
                env.valid_state_check = orig_valid
                rl_mod.policy_forward = orig_pf
            # _info() inside gym.step also calls valid_state_check (differential_gym.py:163):
            # keep every second record (the planner's own call, rrt_rl.py:48)
            rec = rec[1::2]
            p = np.zeros((50, 5)); nz = np.zeros((50, 2)); ac = np.zeros((50, 2))
            for k, (st, _) in enumerate(rec):
                p[k] = st
            for k, z in enumerate(nrec):
                nz[k] = z.reshape(2)
            for k, a in enumerate(arec):
                ac[k] = a
            code = 0
            if not rec[-1][1]:
                code = 1
            elif np.linalg.norm(rec[-1][0][:2] - t[:2]) <= 1.0:
                code = 2
            froms.append(f); targets.append(t); paths.append(p); nsteps.append(len(rec)); status.append(code)
            nnodes.append(len(new)); noises.append(nz); acts.append(ac)
        out[f"{tag}_from"] = np.array(froms)
        out[f"{tag}_target"] = np.array(targets)
        out[f"{tag}_path"] = np.array(paths)
        out[f"{tag}_nsteps"] = np.array(nsteps)
        out[f"{tag}_status"] = np.array(status)
        out[f"{tag}_nnodes"] = np.array(nnodes)
        out[f"{tag}_noise"] = np.array(noises)
        out[f"{tag}_act"] = np.array(acts)
        print("steer_rl", tag, "status", status, "steps", nsteps)
    save("steer_rl", seed=np.array(0), **out)


def g_plan_rl():
    """RRT_RL.planning and RRT_RL_Estimator.planning, seeded random policy, eps=0.05
    with recorded sample and noise streams."""
    maps = load_maps()
    nd = os.path.join(REF, "data", "net", "estimator")
    est, cls = load_estimator(os.path.join(nd, "dist_est_statedict.pth"),
                              os.path.join(nd, "classifier_statedict.pth"))
    out = {}
    for k, (kind, iters) in enumerate((("rl", 25), ("rl_est", 12))):
        env = make_env(maps["test_env1"])
        pol, _ = make_policy(0)
        planner = RRT_RL(env, pol, max_iter=iters) if kind == "rl" else \
            RRT_RL_Estimator(env, pol, est, cls, max_iter=iters)
        planner.set_start_and_goal(np.array([-5, -15, 0, 0, 0.0]), np.array([10, 10, 0, 0, 0.0]))
        rec = tap_samples(planner)
        nrec = []
        orig_randn = torch.randn

        def randn_tap(*a, **kw):
            r = orig_randn(*a, **kw)
            nrec.append(r.numpy().reshape(-1).copy())
            return r
        random.seed(5); np.random.seed(5); torch.manual_seed(5)
        torch.randn = randn_tap
        try:
            with quiet():
                path = planner.planning()
        finally:
            torch.randn = orig_randn
        states, parents, plen = pack_tree(planner)
        out[f"{kind}_samples"] = np.array(rec)
        out[f"{kind}_noise"] = np.array(nrec)
        out[f"{kind}_node_states"] = states
        out[f"{kind}_parents"] = parents
        out[f"{kind}_plen"] = plen
        out[f"{kind}_path"] = np.array(path)
        out[f"{kind}_reach"] = np.array(planner.reach_exactly)
        out[f"{kind}_max_iter"] = np.array(iters)
        print(f"plan_rl {kind}: {len(states)} nodes, {len(rec)} samples, {len(nrec)} control steps")
    save("plan_rl", seed=np.array(0), **out)


def g_gym():
    """DifferentialDriveGym.reset / step / _info / _reward (env/differential_gym.py:136-253):
    seeded episodes of the unmodified reference with recorded velocity commands."""
    maps = load_maps()
    names = [f"obs_list_{i}" for i in range(8)]
    sets = [maps[n] for n in names]
    out = {}
    for ep, (seed, ori) in enumerate(((0, False), (1, False), (2, True), (3, False), (4, False))):
        np.random.seed(seed)
        gym_env = DifferentialDriveGym(make_env(sets[0]), obs_list_list=sets)
        gym_env.set_curriculum(ori=ori)
        obs0 = gym_env.reset()
        picked = [i for i, m in enumerate(sets) if gym_env.robot_env.obs_list is m or gym_env.robot_env.obs_list == m][0]
        rng = np.random.default_rng(100 + seed)
        T = 30 if ep < 4 else 120
        # head toward the goal most of the time so that episodes see goal / collision / free steps
        acts, states, obss, rews, dones, infos = [], [], [], [], [], []
        for t in range(T):
            s, g = gym_env.state, gym_env.goal
            ang = np.arctan2(g[1] - s[1], g[0] - s[0]) - s[2]
            ang = (ang + np.pi) % (2 * np.pi) - np.pi
            a = np.array([rng.uniform(0.3, 1.0), np.clip(2.0 * ang, -np.pi, np.pi) + rng.normal(0, 0.3)])
            if ep == 3:
                a = rng.uniform(gym_env.action_space.low, gym_env.action_space.high)      # random walk
            if ep == 4:
                a = np.array([1.0, 0.0])                                                  # straight into a wall
            obs, rew, done, info = gym_env.step(a)
            acts.append(a); states.append(gym_env.state.copy()); obss.append(obs)
            rews.append(rew); dones.append(done)
            infos.append([float(info[k]) for k in info])
            if done:
                break
        out.update({f"ep{ep}_seed": seed, f"ep{ep}_ori": ori, f"ep{ep}_map": names[picked],
                    f"ep{ep}_start": states[0] * 0 + 0, f"ep{ep}_goal": gym_env.goal.copy(),
                    f"ep{ep}_obs0": obs0, f"ep{ep}_actions": np.array(acts), f"ep{ep}_states": np.array(states),
                    f"ep{ep}_obs": np.array(obss), f"ep{ep}_reward": np.array(rews),
                    f"ep{ep}_done": np.array(dones), f"ep{ep}_info": np.array(infos)})
        # the start state is the state before the first step: recover it by replaying reset
        np.random.seed(seed)
        g2 = DifferentialDriveGym(make_env(sets[0]), obs_list_list=sets)
        g2.set_curriculum(ori=ori)
        g2.reset()
        out[f"ep{ep}_start"] = g2.state.copy()
        assert np.array_equal(g2.goal, gym_env.goal)
        # a2v / v2a round trip values
        out[f"ep{ep}_a2v"] = g2.a2v(np.array(acts[0]))
        out[f"ep{ep}_v2a"] = g2.v2a(np.array(acts[0]))
    out["n_episodes"] = 5
    out["map_names"] = np.array(names)
    save("gym", **out)


def g_circle():
    """CircleRobot footprint (utils/rigid.py:331-347): pairwise dis2obs and whole-map
    validity / clearance with env.rigid_robot = CircleRobot(radius)."""
    from crl_kino.utils.rigid import CircleRobot
    rng = np.random.default_rng(14)
    maps = load_maps()
    out = {}
    radius = 0.6
    robot = CircleRobot(radius)
    poses, rects, dis = [], [], []
    for mname in ("test_env1", "test_env2"):
        for o in maps[mname]:
            l, b, r, t = o.rect.reshape(4)
            for _ in range(30):
                p = np.array([rng.uniform(l - 1.5, r + 1.5), rng.uniform(b - 1.5, t + 1.5), rng.uniform(-np.pi, np.pi)])
                robot.pose = p
                poses.append(p); rects.append(o.rect.reshape(4)); dis.append(robot.dis2obs(o))
    out["radius"] = np.array(radius)
    out["pair_pose"] = np.array(poses)
    out["pair_rect"] = np.array(rects, dtype=np.float32)
    out["pair_dis"] = np.array(dis, dtype=np.float64)
    print("circle pairs:", len(dis), "colliding:", int(np.sum(np.array(dis) == -1)))
    for mname in ("test_env1", "test_env2"):
        env = make_env(maps[mname])
        env.rigid_robot = CircleRobot(radius)
        s = random_states(rng, 200)
        out[f"{mname}_states"] = s
        out[f"{mname}_valid"] = np.array([env.valid_state_check(x) for x in s])
        out[f"{mname}_clear"] = np.array([env.get_clearance(x) for x in s], dtype=np.float64)
    save("circle", **out)


def g_formats():
    """The reference's on-disk formats: bytes pickled BY THE REFERENCE's classes
    (obstacle list, list of obstacle lists, node_list, path, benchmark positions, results dict)
    plus the arrays they hold; and the converse check, done here once: files written by
    crl_kino_b200.utils.io load under the unmodified reference."""
    maps = load_maps()
    out = {}
    as_u8 = lambda b: np.frombuffer(b, dtype=np.uint8)
    out["obs_pickle"] = as_u8(pickle.dumps(maps["test_env2"]))
    out["obs_rects"] = rects_of(maps["test_env2"])
    lol = [maps["obs_list_0"], maps["obs_list_1"]]
    out["obs_lol_pickle"] = as_u8(pickle.dumps(lol))
    out["obs_lol_rects0"] = rects_of(lol[0])
    out["obs_lol_rects1"] = rects_of(lol[1])
    pos = pickle.load(open(os.path.join(REF, "examples", "benchmark_position_test_env2.pkl"), "rb"))
    out["positions_pickle"] = as_u8(pickle.dumps(pos))
    out["positions"] = np.array(pos, dtype=np.float64)
    env = make_env(maps["test_env1"])
    planner = RRT(env, max_iter=4)
    planner.set_start_and_goal(np.array([-5, -15, 0, 0, 0.0]), np.array([10, 10, 0, 0, 0.0]))
    random.seed(5); np.random.seed(5)
    with quiet():
        path = planner.planning()
    states, parents, plen = pack_tree(planner)
    out["tree_pickle"] = as_u8(pickle.dumps(planner.node_list))
    out["tree_states"] = states
    out["tree_parents"] = parents
    out["tree_plen"] = plen
    out["path_pickle"] = as_u8(pickle.dumps(path))
    out["path"] = np.array(path)
    data = {"runtime": np.array([1.25, -1, 3.5]), "path_len": np.array([12.4, 30.0])}     # benchmark.py:86-88
    out["results_pickle"] = as_u8(pickle.dumps(data))
    # converse direction: our writers, the reference's loaders
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from crl_kino_b200.utils import io as cio
    back = pickle.loads(cio.save_obstacles(out["obs_rects"]))
    assert type(back[0]) is RectObs and np.array_equal(rects_of(back), out["obs_rects"])
    assert back[3].point_in_obstacle(back[3].rect.mean(0))
    nodes = pickle.loads(cio.save_tree(planner.node_list))
    assert type(nodes[0]) is RRT.Node
    s2, p2, l2 = pack_tree(type("P", (), {"node_list": nodes})())
    assert np.array_equal(s2, states) and np.array_equal(p2, parents) and np.array_equal(l2, plen)
    print("formats: reference loaded the files written by crl_kino_b200.utils.io")
    save("formats", **out)


# ------------------------------------------------- contract-size runs (BASELINE configs 1 and 3)

def _plan_dwa_one(a):
    """One full RRT.planning run of the unmodified reference (worker of a process pool)."""
    mname, start, goal, seed, iters = a
    maps = load_maps()
    env = make_env(maps[mname])
    planner = RRT(env, max_iter=iters)
    planner.set_start_and_goal(np.array(start, dtype=float), np.array(goal, dtype=float))
    rec = tap_samples(planner)
    random.seed(seed); np.random.seed(seed)
    steps = {"n": 0, "ext": 0}
    oc = planner.dwa.control
    ost = planner.steer

    def ctrl(x, g, _s=steps, _oc=oc):
        _s["n"] += 1
        return _oc(x, g)

    def steer(*a, _s=steps, **k):
        _s["ext"] += 1
        return ost(*a, **k)
    planner.dwa.control = ctrl
    planner.steer = steer
    tic = time.time()
    with quiet():
        path = planner.planning()
    el = time.time() - tic
    states, parents, plen = pack_tree(planner)
    return dict(samples=np.array(rec), node_states=states, parents=parents, plen=plen, path=np.array(path),
                reach=np.array(planner.reach_exactly), ctrl_steps=np.array(steps["n"]),
                extensions=np.array(steps["ext"]), ref_seconds=np.array(el),
                planning_time=np.array(planner.planning_time))


def _pool_runs(jobs, workers):
    import multiprocessing as mp
    with mp.get_context("fork").Pool(workers) as pool:
        return pool.map(_plan_dwa_one, jobs, chunksize=1)


def g_plan_dwa_full():
    """BASELINE config 1 at contract size (examples/test.py:144-160): test_env1, start
    [-5,-15,0,0,0] -> goal [10,10,0,0,0], RRT defaults (max_iter 500, rrt.py:35), seeds 0..9."""
    jobs = [("test_env1", [-5, -15, 0, 0, 0], [10, 10, 0, 0, 0], s, 500) for s in range(10)]
    res = _pool_runs(jobs, int(os.environ.get("GOLDEN_WORKERS", "5")))
    out = {"n_runs": np.array(len(jobs)), "max_iter": np.array(500)}
    for k, (job, r) in enumerate(zip(jobs, res)):
        out[f"run{k}_seed"] = np.array(job[3])
        out[f"run{k}_start"] = np.array(job[1], dtype=float)
        out[f"run{k}_goal"] = np.array(job[2], dtype=float)
        for key, v in r.items():
            out[f"run{k}_{key}"] = v
        print(f"plan_dwa_full seed {job[3]}: {len(r['node_states'])} nodes, {len(r['samples'])} samples, "
              f"{int(r['extensions'])} extensions, {int(r['ctrl_steps'])} control steps, reach={bool(r['reach'])}, "
              f"{float(r['ref_seconds']):.1f}s")
    save("plan_dwa_full", **out)


def g_plan_cfg3():
    """BASELINE config 3 (examples/benchmark.py:56-75) with the reference's own RRT (DWA steering):
    all 12 ordered (start, goal) pairs of benchmark_position_test_env{1,2}.pkl on test_env{1,2}.pkl,
    max_iter 500, seed = 100 * env + pair index."""
    jobs, meta = [], []
    for e, mname in ((1, "test_env1"), (2, "test_env2")):
        pos = pickle.load(open(os.path.join(REF, "examples", f"benchmark_position_test_env{e}.pkl"), "rb"))
        pos = np.asarray(pos, dtype=float)
        k = 0
        for i in range(len(pos)):
            for j in range(len(pos)):
                if i == j:
                    continue
                jobs.append((mname, pos[i], pos[j], 100 * e + k, 500))
                meta.append((e, i, j))
                k += 1
    res = _pool_runs(jobs, int(os.environ.get("GOLDEN_WORKERS", "5")))
    out = {"n_runs": np.array(len(jobs)), "max_iter": np.array(500)}
    for k, (job, m, r) in enumerate(zip(jobs, meta, res)):
        out[f"run{k}_env"] = np.array(m[0]); out[f"run{k}_i"] = np.array(m[1]); out[f"run{k}_j"] = np.array(m[2])
        out[f"run{k}_seed"] = np.array(job[3])
        out[f"run{k}_start"] = np.array(job[1], dtype=float)
        out[f"run{k}_goal"] = np.array(job[2], dtype=float)
        for key, v in r.items():
            out[f"run{k}_{key}"] = v
        # benchmark.py:69-75
        out[f"run{k}_runtime"] = np.array(0.2 * (len(r["path"]) - 1) if bool(r["reach"]) else -1.0)
        print(f"plan_cfg3 env{m[0]} {m[1]}->{m[2]}: {len(r['node_states'])} nodes, {len(r['samples'])} samples, "
              f"{int(r['extensions'])} extensions, reach={bool(r['reach'])}, {float(r['ref_seconds']):.1f}s")
    save("plan_cfg3", **out)


def _steer_rl_dense_one(a):
    """One RRT_RL.steer of the unmodified reference on the cfg4 map (worker of a process pool)."""
    rects, f, t, seed = a
    import crl_kino.planner.rrt_rl as rl_mod
    env = make_env([RectObs(r.reshape(2, 2)) for r in rects])
    pol, actor = make_policy(seed)
    planner = RRT_RL(env, pol)
    rec, arec = [], []
    orig_pf = rl_mod.policy_forward

    def pf(policy, obs, info=None, eps=0.0):
        a_ = orig_pf(policy, obs, info=info, eps=0.0)
        arec.append(a_[0].copy())
        return a_
    rl_mod.policy_forward = pf
    orig_valid = env.valid_state_check

    def tap(state, _rec=rec, _o=orig_valid):
        r = _o(state)
        _rec.append((np.array(state), r))
        return r
    env.valid_state_check = tap
    planner.node_list = []
    tic = time.time()
    new = planner.steer(planner.Node(np.array(f)), planner.Node(np.array(t)))
    el = time.time() - tic
    rec = rec[1::2]          # _info() inside gym.step also calls valid_state_check (differential_gym.py:163)
    p = np.zeros((50, 5)); ac = np.zeros((50, 2))
    for k, (st, _) in enumerate(rec):
        p[k] = st
    for k, a_ in enumerate(arec):
        ac[k] = a_
    code = 0
    if not rec[-1][1]:
        code = 1
    elif np.linalg.norm(rec[-1][0][:2] - np.array(t)[:2]) <= 1.0:
        code = 2
    return dict(path=p, act=ac, nsteps=len(rec), status=code, nnodes=len(new), seconds=el)


def g_steer_rl_dense():
    """RRT_RL.steer (rrt_rl.py:22-63) of the unmodified reference on BASELINE config 4's map (10,004
    obstacles, crl_kino_b200.workloads.dense_map), seeded random-init policy, eps = 0: start states with
    no obstacle within 1.5 m (conservative raster), targets 3-20 m away.  ~1.5 s of reference time per
    policy step (valid_state_check and get_clearance over every obstacle in Python)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from crl_kino_b200.workloads import clearance_raster, dense_map, raster_free
    import multiprocessing as mp
    rects = dense_map()
    rng = np.random.default_rng(21)
    gx, occ = clearance_raster(rects, 1.5)
    froms, targets = [], []
    while len(froms) < 8:
        f = random_states(rng, 1)[0]
        f[3:] = 0
        if not raster_free(gx, occ, f[None])[0]:
            continue
        t = random_states(rng, 1)[0]
        if not (3.0 <= np.linalg.norm(f[:2] - t[:2]) < 20):
            continue
        froms.append(f); targets.append(t)
    with mp.get_context("fork").Pool(int(os.environ.get("GOLDEN_WORKERS", "4"))) as pool:
        res = pool.map(_steer_rl_dense_one, [(rects, froms[i], targets[i], 0) for i in range(len(froms))], chunksize=1)
    out = {"seed": np.array(0), "n_cells": np.array(10000), "from": np.array(froms), "target": np.array(targets)}
    for key in ("path", "act", "nsteps", "status", "nnodes", "seconds"):
        out[key] = np.array([r[key] for r in res])
    for i, r in enumerate(res):
        print(f"steer_rl_dense {i}: {r['nsteps']} steps, status {r['status']}, {r['nnodes']} nodes, {r['seconds']:.0f}s")
    save("steer_rl_dense", **out)


ALL = dict(fixtures=g_fixtures, numerics=g_numerics, motion=g_motion, dwa=g_dwa,
           dwa_clearance=g_dwa_clearance, geometry=g_geometry, obs=g_obs, nearest=g_nearest,
           steer_dwa=g_steer_dwa, plan_dwa=g_plan_dwa, policy=g_policy, estimator=g_estimator,
           steer_rl=g_steer_rl, plan_rl=g_plan_rl, gym=g_gym, circle=g_circle, formats=g_formats,
           plan_dwa_full=g_plan_dwa_full, plan_cfg3=g_plan_cfg3, steer_rl_dense=g_steer_rl_dense)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    a = ap.parse_args()
    for name, fn in ALL.items():
        if a.only and name not in a.only:
            continue
        t0 = time.time()
        fn()
        print(f"[{name}] {time.time()-t0:.1f}s")
