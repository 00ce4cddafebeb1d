"""ctypes front-end of the CPU oracle (oracle/crlk_oracle.c) plus the float32
network restatements and the planner loop.

TEST INFRASTRUCTURE ONLY -- see the header of crlk_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module; the product package never does.

Parity status: PINNED against tests/golden/*.npz (vectors generated from the
unmodified reference by tests/golden/make_golden.py); see
tests/test_oracle_golden.py.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcrlk_oracle.so")


def build(force=False):
    """Compile crlk_oracle.c with gcc (the recipe is oracle/Makefile)."""
    src = os.path.join(_HERE, "crlk_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


class OrcEnv(C.Structure):
    _fields_ = [("max_v", C.c_double), ("min_v", C.c_double), ("max_w", C.c_double),
                ("max_acc_v", C.c_double), ("max_acc_w", C.c_double), ("base_dt", C.c_double),
                ("env_bounds", C.c_double * 4), ("robot_rect", C.c_float * 4), ("robot_radius", C.c_double)]


class OrcDwa(C.Structure):
    _fields_ = [("v_res", C.c_double), ("w_res", C.c_double), ("dt", C.c_double),
                ("predict_time", C.c_double), ("to_goal_cost_gain", C.c_double),
                ("ob_gain", C.c_double), ("speed_cost_gain", C.c_double),
                ("skip", C.c_int), ("use_clearance", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_normalize_angle.restype = C.c_double
        _lib.orc_normalize_angle.argtypes = [C.c_double]
        _lib.orc_pairwise_sum.restype = C.c_double
        _lib.orc_pairwise_sum.argtypes = [C.c_void_p, C.c_long, C.c_long]
        _lib.orc_dis2obs.restype = C.c_double
        _lib.orc_dis2obs_circle.restype = C.c_double
        _lib.orc_dis2obs_circle.argtypes = [C.c_double, C.c_void_p, C.c_void_p]
        _lib.orc_get_clearance.restype = C.c_double
        _lib.orc_arange.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int]
        _lib.orc_motion_velocity.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        _lib.orc_motion.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        _lib.orc_calc_trajectory.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                             C.c_void_p, C.c_int]
        _lib.orc_dynamic_window.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        _lib.orc_nearest.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_double, C.c_double,
                                     C.c_void_p, C.c_void_p]
        _lib.orc_nearest_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p,
                                           C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


ROBOT_RECT = np.array([-0.7, -0.5, 0.7, 0.5], dtype=np.float32)  # differential_env.py:41


def cpu_count():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class Env:
    """Oracle counterpart of DifferentialDriveEnv (differential_env.py:27-169)."""

    def __init__(self, max_v=1.0, min_v=-0.1, max_w=math.pi, max_acc_v=1.0, max_acc_w=math.pi,
                 env_bounds=(-20, 20, -20, 20), base_dt=0.1, robot_rect=ROBOT_RECT, robot_radius=0.0):
        """robot_radius > 0: CircleRobot(radius) footprint (rigid.py:331-347) instead of the rectangle."""
        self.p = OrcEnv(max_v, min_v, max_w, max_acc_v, max_acc_w, base_dt,
                        (C.c_double * 4)(*[float(x) for x in env_bounds]),
                        (C.c_float * 4)(*[float(x) for x in robot_rect]), float(robot_radius))
        self.rects = np.zeros((0, 4), dtype=np.float32)
        self.max_v, self.min_v, self.max_w = max_v, min_v, max_w
        self.max_acc_v, self.max_acc_w, self.base_dt = max_acc_v, max_acc_w, base_dt
        self.env_bounds = np.array(env_bounds, dtype=np.float64).reshape(2, 2)

    @property
    def ref(self):
        return C.byref(self.p)

    def set_obs(self, rects):
        """rects: M x 4 float32 [left, bottom, right, top]."""
        self.rects = np.ascontiguousarray(np.asarray(rects, dtype=np.float32).reshape(-1, 4))

    @property
    def M(self):
        return len(self.rects)

    def get_bounds(self):
        sb = np.zeros((5, 2))
        sb[:2] = self.env_bounds
        sb[2] = [-np.pi, np.pi]
        sb[3] = [self.min_v, self.max_v]
        sb[4] = [-self.max_w, self.max_w]
        return sb

    def dynamic_window(self, state, dt):
        out = np.zeros(4)
        lib().orc_dynamic_window(self.ref, _p(_f64(state)), dt, _p(out))
        return out.reshape(2, 2)

    def motion_velocity(self, state, cmd, duration):
        out = np.zeros(5)
        lib().orc_motion_velocity(self.ref, _p(_f64(state)), _p(_f64(cmd)), float(duration), _p(out))
        return out

    def motion(self, state, acc, duration):
        out = np.zeros(5)
        lib().orc_motion(self.ref, _p(_f64(state)), _p(_f64(acc)), float(duration), _p(out))
        return out

    def calc_trajectory(self, x_init, cmd, dt, predict_time):
        traj = np.zeros((64, 5))
        rows = lib().orc_calc_trajectory(self.ref, _p(_f64(x_init)), _p(_f64(cmd)), float(dt),
                                         float(predict_time), _p(traj), 64)
        return traj[:rows].copy()

    def dis2obs(self, pose, rect):
        rect = np.ascontiguousarray(rect, dtype=np.float32).reshape(4)
        rr = np.ctypeslib.as_array(self.p.robot_rect).astype(np.float32)
        return lib().orc_dis2obs(_p(rr), _p(_f64(pose)), _p(rect))

    def dis2obs_circle(self, radius, pose, rect):
        rect = np.ascontiguousarray(rect, dtype=np.float32).reshape(4)
        return lib().orc_dis2obs_circle(float(radius), _p(_f64(pose)), _p(rect))

    def valid_state_check(self, state):
        return bool(lib().orc_valid_state(self.ref, _p(self.rects), self.M, _p(_f64(state))))

    def get_clearance(self, state):
        return lib().orc_get_clearance(self.ref, _p(self.rects), self.M, _p(_f64(state)))

    def valid_state_batch(self, states, clearance=True, threads=None):
        states = _f64(states).reshape(-1, 5)
        B = len(states)
        valid = np.zeros(B, dtype=np.uint8)
        clr = np.zeros(B) if clearance else None
        lib().orc_valid_state_batch(self.ref, _p(self.rects), self.M, _p(states), B, _p(valid),
                                    _p(clr), threads or cpu_count())
        return valid.astype(bool), clr

    def valid_point_check(self, pts):
        pts = _f64(pts).reshape(-1, 2)
        out = np.zeros(len(pts), dtype=np.uint8)
        lib().orc_valid_points(_p(self.rects), self.M, _p(pts), len(pts), _p(out))
        return out.astype(bool)

    def local_obs(self, state, goal):
        s = sample_positions()
        out = np.zeros(4 + len(s))
        lib().orc_local_obs(_p(self.rects), self.M, _p(_f64(state)), _p(_f64(goal)[:2].copy()),
                            _p(s), len(s), _p(out))
        return out

    def local_obs_batch(self, states, goals, threads=None):
        states = _f64(states).reshape(-1, 5)
        goals = np.ascontiguousarray(_f64(goals).reshape(len(states), -1)[:, :2])
        s = sample_positions()
        out = np.zeros((len(states), 4 + len(s)))
        lib().orc_local_obs_batch(_p(self.rects), self.M, _p(states), _p(goals), _p(s), len(s),
                                  len(states), _p(out), threads or cpu_count())
        return out


def arange(start, stop, step):
    cap = 1 << 16
    out = np.zeros(cap)
    n = lib().orc_arange(float(start), float(stop), float(step), _p(out), cap)
    return out[:n].copy()


def pairwise_sum(a, stride=1):
    a = _f64(a)
    return lib().orc_pairwise_sum(_p(a), len(a) // stride if stride > 1 else len(a), stride)


_SAMPLES = None


def sample_positions():
    """differential_gym.py:62-75: product(ba, lr), forward axis first."""
    global _SAMPLES
    if _SAMPLES is None:
        lr = arange(-4, 4, 0.3)
        ba = arange(-2, 6, 0.3)
        _SAMPLES = np.ascontiguousarray(
            np.array([(b, l) for b in ba for l in lr], dtype=np.float64))
    return _SAMPLES


class Dwa:
    """Oracle counterpart of DWA (dwa.py:12-86)."""

    def __init__(self, env, v_res=0.01, w_res=0.1, dt=0.1, predict_time=1.0, to_goal_cost_gain=1.0,
                 ob_gain=1.0, speed_cost_gain=0.3, skip=3, use_clearance=False):
        self.env = env
        self.p = OrcDwa(v_res, w_res, dt, predict_time, to_goal_cost_gain, ob_gain, speed_cost_gain,
                        skip, int(use_clearance))

    def set_dwa(self, **kw):
        for k, v in kw.items():
            setattr(self.p, k, v)

    def control(self, x_init, goal, want_costs=False):
        e = self.env
        opt_v = np.zeros(2)
        idx = C.c_int(0)
        cost = C.c_double(0)
        traj = np.zeros((64, 5))
        rows = C.c_int(0)
        cap = 1 << 20
        cs = np.zeros(cap) if want_costs else None
        R = lib().orc_dwa_control(e.ref, C.byref(self.p), _p(e.rects), e.M, _p(_f64(x_init)),
                                  _p(_f64(goal)[:2].copy()), _p(opt_v), C.byref(idx), C.byref(cost),
                                  _p(traj), 64, C.byref(rows), _p(cs), cap)
        assert R >= 0
        res = dict(opt_v=opt_v, opt_idx=idx.value, opt_cost=cost.value, opt_traj=traj[:rows.value].copy(),
                   R=R)
        if want_costs:
            res["cost_sum"] = cs[:R].copy()
        return res

    def control_batch(self, states, goals, threads=None):
        e = self.env
        states = _f64(states).reshape(-1, 5)
        B = len(states)
        goals = np.ascontiguousarray(_f64(goals).reshape(B, -1)[:, :2])
        opt_v = np.zeros((B, 2))
        idx = np.zeros(B, dtype=np.int32)
        cost = np.zeros(B)
        R = np.zeros(B, dtype=np.int32)
        lib().orc_dwa_control_batch(e.ref, C.byref(self.p), _p(e.rects), e.M, _p(states), _p(goals), B,
                                    _p(opt_v), _p(idx), _p(cost), _p(R), threads or cpu_count())
        return dict(opt_v=opt_v, opt_idx=idx, opt_cost=cost, R=R)

    def extend_batch(self, from_states, targets, n_max_extend=50, n_tree=25, threads=None,
                     want_ctrl=False):
        """rrt.py:99-135 for Q independent (from, target) pairs."""
        e = self.env
        from_states = _f64(from_states).reshape(-1, 5)
        Q = len(from_states)
        targets = _f64(targets).reshape(Q, 5)
        per = n_max_extend // n_tree + 1
        paths = np.zeros((Q, n_max_extend, 5))
        n_steps = np.zeros(Q, dtype=np.int32)
        status = np.zeros(Q, dtype=np.int32)
        node_steps = np.zeros((Q, per), dtype=np.int32)
        n_nodes = np.zeros(Q, dtype=np.int32)
        ctrl = np.full((Q, n_max_extend), -1, dtype=np.int32) if want_ctrl else None
        lib().orc_extend_dwa_batch(e.ref, C.byref(self.p), _p(e.rects), e.M, _p(from_states),
                                   _p(targets), Q, n_max_extend, n_tree, _p(paths), _p(n_steps),
                                   _p(status), _p(node_steps), _p(n_nodes), _p(ctrl),
                                   threads or cpu_count())
        return dict(path=paths, n_steps=n_steps, status=status, node_step=node_steps, n_nodes=n_nodes,
                    ctrl_idx=ctrl)


def nearest(xs, ys, target):
    xs, ys = _f64(xs), _f64(ys)
    idx = C.c_int(0)
    d = C.c_double(0)
    lib().orc_nearest(_p(xs), _p(ys), len(xs), float(target[0]), float(target[1]), C.byref(idx),
                      C.byref(d))
    return idx.value, d.value


def nearest_batch(xs, ys, n_nodes, targets, threads=None):
    """xs, ys: Q x stride SoA trees; n_nodes: Q; targets: Q x 2."""
    xs, ys = _f64(xs), _f64(ys)
    Q, stride = xs.shape
    n_nodes = np.ascontiguousarray(n_nodes, dtype=np.int32)
    targets = np.ascontiguousarray(_f64(targets).reshape(Q, -1)[:, :2])
    idx = np.zeros(Q, dtype=np.int32)
    dist = np.zeros(Q)
    lib().orc_nearest_batch(_p(xs), _p(ys), _p(n_nodes), stride, _p(targets), Q, _p(idx), _p(dist),
                            threads or cpu_count())
    return idx, dist


# ------------------------------------------------------------------ networks

ACTION_LOW = np.array([-0.1, -np.pi], dtype=np.float32)   # gym Box float32 (differential_gym.py:41-44)
ACTION_HIGH = np.array([1.0, np.pi], dtype=np.float32)


def actor_forward(weights, biases, obs, noise=None, eps=0.0, low=ACTION_LOW, high=ACTION_HIGH):
    """policy/net.py:140-154 + tianshou ddpg.py:101-131, float32 throughout.

    weights[i]: [out, in] float32 (torch nn.Linear layout); noise: B x 2 standard
    normal draws (the reference draws torch.randn inside the actor).
    """
    x = np.asarray(obs).astype(np.float32).reshape(-1, weights[0].shape[1])
    n = len(weights)
    for i, (W, b) in enumerate(zip(weights, biases)):
        x = (x @ W.T.astype(np.float32) + b.astype(np.float32)).astype(np.float32)
        if i < n - 1:
            x = np.maximum(x, np.float32(0))
    x = np.tanh(x).astype(np.float32)
    if noise is not None:
        x = (x + np.asarray(noise, dtype=np.float32) * np.float32(eps)).astype(np.float32)
    bias = ((low + high) / np.float32(2)).astype(np.float32)
    scale = ((high - low) / np.float32(2)).astype(np.float32)
    x = (x * scale[None] + bias[None]).astype(np.float32)
    x = np.minimum(np.maximum(x, low[None]), high[None])
    x = np.minimum(np.maximum(x, low[None]), high[None])
    return x


def cnn_forward(sd, spatial, ext, sigmoid):
    """estimator/network.py:39-69 / :97-127 in torch CPU float32 (the oracle
    of a floating-point kernel may be a torch fp32 restatement)."""
    import torch
    import torch.nn.functional as F
    with torch.no_grad():
        x = torch.as_tensor(np.asarray(spatial, dtype=np.float32))
        e = torch.as_tensor(np.asarray(ext, dtype=np.float32))
        t = {k: torch.as_tensor(np.asarray(v, dtype=np.float32)) for k, v in sd.items()}
        for i in (1, 2, 3):
            x = F.max_pool2d(F.relu(F.conv2d(x, t[f"conv{i}.weight"], t[f"conv{i}.bias"], padding=1)), 2)
        x = F.relu(F.conv2d(x, t["conv4.weight"], t["conv4.bias"], padding=1))
        x = F.adaptive_avg_pool2d(x, (1, 1)).flatten(1)
        y = F.linear(torch.cat([x, e], 1), t["linear.weight"], t["linear.bias"])
        if sigmoid:
            y = torch.sigmoid(y)
        return y.numpy()


def estimator_inputs(obs):
    """rrt_rl_estimator.py:57-62: obs[B,733] -> spatial[B,1,27,27], ext[B,3]."""
    inp = np.asarray(obs).astype(np.float32).copy()
    import torch
    inp[:, 1] = torch.norm(torch.as_tensor(inp[:, :2]), dim=1).numpy()
    return inp[:, 4:].reshape(-1, 1, 27, 27), inp[:, 1:4]


# ------------------------------------------------------------------- planner

class Node:
    __slots__ = ("state", "path", "parent")

    def __init__(self, state):
        self.state = np.array(state, dtype=np.float64)
        self.path = []
        self.parent = None


class PlannerPort:
    """rrt.py:21-161 / rrt_rl.py / rrt_rl_estimator.py driven by an explicit
    sample stream (states the reference's `sample` would have returned, in
    order) instead of Python's global RNGs.

    steering: 'dwa' | 'rl' ; parent: 'nearest' | 'estimator'.
    """

    def __init__(self, env, steering="dwa", actor=None, estimator=None, classifier=None,
                 max_iter=500, eps=0.05, noise_stream=None, dwa_kwargs=None):
        self.env = env
        self.dwa = Dwa(env, **(dwa_kwargs or dict(dt=0.2, v_res=0.02)))
        self.steering = steering
        self.actor = actor            # (weights, biases)
        self.estimator = estimator    # state dict of numpy arrays
        self.classifier = classifier
        self.max_iter = max_iter
        self.eps = eps
        self.noise_stream = noise_stream  # iterator of [2] float32 draws
        self.node_list = []
        self.n_extend_calls = 0
        self.n_control_steps = 0
        self.ctrl_log = []

    def set_start_and_goal(self, start, goal):
        assert self.env.valid_state_check(start) and self.env.valid_state_check(goal)
        self.start = Node(start)
        self.goal = Node(goal)

    def choose_parent(self, node_list, target):
        xs = np.array([n.state[0] for n in node_list])
        ys = np.array([n.state[1] for n in node_list])
        if self.estimator is None:
            return nearest(xs, ys, target.state)
        # rrt_rl_estimator.py:26-79
        d = np.array([nearest([x], [y], target.state)[1] for x, y in zip(xs, ys)])
        valid = np.where((d >= 1.0) & (d < 20.0))[0]
        if len(valid) == 0:
            return 0, 100
        obs = self.env.local_obs_batch(np.array([node_list[i].state for i in valid]),
                                       np.tile(target.state, (len(valid), 1)), threads=1)
        spatial, ext = estimator_inputs(obs)
        est = cnn_forward(self.estimator, spatial, ext, False)[:, 0]
        prob = cnn_forward(self.classifier, spatial, ext, True)[:, 0]
        fail = 1 - np.round(prob)
        d[valid] = (est + 1) * 15
        d[valid] += 100 * fail
        i = int(np.argmin(d))
        return i, d[i]

    def steer(self, from_node, to_node, t_max_extend=10.0, t_tree=5.0):
        self.n_extend_calls += 1
        parent = from_node
        new = Node(from_node.state)
        new.parent = parent
        if self.steering != "dwa":
            new.path.append(new.state)      # rrt_rl.py:27
        state = new.state
        goal = to_node.state.copy()
        dt = 0.2
        n_max, n_tree = round(t_max_extend / dt), round(t_tree / dt)
        out = []
        for k in range(1, n_max + 1):
            self.n_control_steps += 1
            if self.steering == "dwa":
                r = self.dwa.control(state, goal)
                v = r["opt_v"]
                self.ctrl_log.append(r["opt_idx"])
            else:
                obs = self.env.local_obs(state, goal)
                noise = None
                if self.noise_stream is not None:
                    noise = np.asarray(next(self.noise_stream), dtype=np.float32).reshape(1, 2)
                v = actor_forward(self.actor[0], self.actor[1], obs, noise, self.eps)[0].astype(np.float64)
            state = self.env.motion_velocity(state, v, dt)
            new.state = state
            new.path.append(state)
            if not self.env.valid_state_check(state):
                break
            elif nearest([state[0]], [state[1]], goal)[1] <= 1.0:
                out.append(new)
                break
            elif k % n_tree == 0:
                out.append(new)
                parent = new
                new = Node(parent.state)
                new.parent = parent
                new.path.append(new.state)
        self.node_list += out
        return out

    def planning(self, sample_stream):
        """sample_stream: iterator over 5-vectors (every draw of rrt.py:146-154)."""
        self.node_list = [self.start]
        self.n_samples = 0
        reach = False
        gxy = self.goal.state
        for _ in range(self.max_iter):
            while True:
                rnd = Node(next(sample_stream))
                self.n_samples += 1
                if not self.env.valid_state_check(rnd.state):
                    continue
                ind, cost = self.choose_parent(self.node_list, rnd)
                if 1.0 <= cost < 20:
                    break
            new_nodes = self.steer(self.node_list[ind], rnd)
            if len(new_nodes) > 0:
                last = self.node_list[-1].state
                if nearest([last[0]], [last[1]], gxy)[1] <= 1.0:
                    reach = True
                    break
                ind, cost = self.choose_parent(new_nodes, self.goal)
                if cost < 5.0:
                    self.steer(new_nodes[ind], self.goal)
                    last = self.node_list[-1].state
                    if nearest([last[0]], [last[1]], gxy)[1] <= 1.0:
                        reach = True
                        break
        self.reach_exactly = reach
        self.path = self.generate_final_course(-1)
        return self.path

    def generate_final_course(self, goal_ind):
        path = []
        node = self.node_list[goal_ind]
        while node.parent is not None:
            path = node.path[1:] + path
            node = node.parent
        return [node.state] + path


# ------------------------------------------------------------------ gym side
# DifferentialDriveGym.step / _info / _reward (env/differential_gym.py:136-178)

GYM_INFO_KEYS = ("goal", "goal_dis", "heading", "collision", "clearance", "v", "w", "step")
GYM_REWARD_PARAM = np.array([50.0, -1.5, -0.5, -300.0, 1.0, 1.0, -.5, -2.5])   # differential_gym.py:28-29


def _normalize_angle(a):
    """differential_env.py:20-24."""
    m = a % (2 * math.pi)
    if m > math.pi:
        m -= 2 * math.pi
    return m


def gym_info(env, state, goal):
    """differential_gym.py:147-166 as the 8-vector in dict order."""
    state = np.asarray(state, dtype=np.float64)
    goal = np.asarray(goal, dtype=np.float64)
    gd = float(np.linalg.norm(state[:2] - goal[:2]))
    head = abs(_normalize_angle(np.arctan2(goal[1] - state[1], goal[0] - state[0]) - state[2]))
    clr = min(env.get_clearance(state), 1.0)
    col = not env.valid_state_check(state)
    return np.array([float(gd <= 1.0), gd, head, float(col), clr, state[3], abs(state[4]), 1.0])


def gym_step(env, state, goal, action, reward_param=GYM_REWARD_PARAM, dt=1.0 / 5.0):
    """differential_gym.py:136-145 -> (state', obs, reward, info vector)."""
    s = env.motion_velocity(state, action, dt)
    info = gym_info(env, s, goal)
    return s, env.local_obs(s, goal), float(info @ np.asarray(reward_param, dtype=np.float64)), info
