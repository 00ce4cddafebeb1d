"""Device-level operator layer: thin, allocation-explicit wrappers over the C ABI
(include/crlk.h) that take and return torch CUDA tensors.  PyTorch is plumbing
only (device memory, streams); every computation happens in libcrlk.so."""
import ctypes as C
import math

import numpy as np
import torch

from ._lib import DwaParams, EnvParams, ExtendParams, CrlkError, check, lib, require_cuda

ROBOT_RECT = (-0.7, -0.5, 0.7, 0.5)   # differential_env.py:41
OBS_DIM = 733
OBS_LD = 736                          # obs row stride for the MLP (16-byte aligned rows)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dev():
    return torch.device("cuda", torch.cuda.current_device())


def as_dev(a, dtype=torch.float64):
    """numpy / list / tensor -> contiguous CUDA tensor of dtype."""
    if isinstance(a, torch.Tensor):
        return a.to(device=dev(), dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(dev())


class EnvHandle:
    """Owns a CrlkEnv* (obstacle set + vehicle limits on the device)."""

    def __init__(self, max_v, min_v, max_w, max_acc_v, max_acc_w, base_dt, env_bounds, rects,
                 robot_rect=ROBOT_RECT, robot_radius=0.0):
        require_cuda()
        self.params = EnvParams(max_v, min_v, max_w, max_acc_v, max_acc_w, base_dt,
                                (C.c_double * 4)(*[float(x) for x in env_bounds]),
                                (C.c_float * 4)(*[float(x) for x in robot_rect]), float(robot_radius))
        rects = np.ascontiguousarray(np.asarray(rects, dtype=np.float32).reshape(-1, 4))
        self.M = len(rects)
        h = C.c_void_p()
        check(lib().crlk_env_create(C.byref(self.params), rects.ctypes.data_as(C.c_void_p), self.M,
                                    C.byref(h)))
        self.h = h
        self.device = torch.cuda.current_device()

    def set_obs(self, rects):
        rects = np.ascontiguousarray(np.asarray(rects, dtype=np.float32).reshape(-1, 4))
        self.M = len(rects)
        check(lib().crlk_env_set_obs(self.h, rects.ctypes.data_as(C.c_void_p), self.M))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().crlk_env_destroy(self.h)
                self.h = None
        except Exception:
            pass


def dwa_params(v_res, w_res, dt, predict_time, to_goal_cost_gain, ob_gain, speed_cost_gain, skip,
               use_clearance=False):
    return DwaParams(v_res, w_res, dt, predict_time, to_goal_cost_gain, ob_gain, speed_cost_gain,
                     int(skip), int(bool(use_clearance)))


def extend_params(t_max_extend=10.0, t_tree=5.0, dt=0.2, reach_radius=1.0):
    """rrt.py:109-111."""
    return ExtendParams(round(t_max_extend / dt), round(t_tree / dt), dt, reach_radius)


def valid_state(env, states, want_clearance=False, want_near=False):
    states = as_dev(states).reshape(-1, 5)
    B = states.shape[0]
    valid = torch.empty(B, dtype=torch.uint8, device=states.device)
    clr = torch.empty(B, dtype=torch.float64, device=states.device) if want_clearance else None
    near = torch.empty(B, dtype=torch.uint8, device=states.device) if want_near else None
    check(lib().crlk_valid_state(env.h, _ptr(states), B, _ptr(valid), _ptr(clr), _ptr(near), _stream()))
    return valid, clr, near


def valid_points(env, points):
    points = as_dev(points).reshape(-1, 2)
    K = points.shape[0]
    valid = torch.empty(K, dtype=torch.uint8, device=points.device)
    check(lib().crlk_valid_points(env.h, _ptr(points), K, _ptr(valid), _stream()))
    return valid


def motion_velocity(env, states, cmds, duration, accel=False):
    states = as_dev(states).reshape(-1, 5)
    cmds = as_dev(cmds).reshape(-1, 2)
    out = torch.empty_like(states)
    fn = lib().crlk_motion if accel else lib().crlk_motion_velocity
    check(fn(env.h, _ptr(states), _ptr(cmds), float(duration), states.shape[0], _ptr(out), _stream()))
    return out


def dynamic_window(env, states, dt):
    states = as_dev(states).reshape(-1, 5)
    out = torch.empty(states.shape[0], 4, dtype=torch.float64, device=states.device)
    check(lib().crlk_dynamic_window(env.h, _ptr(states), float(dt), states.shape[0], _ptr(out), _stream()))
    return out


def local_obs(env, states, goals, want64=True, want32=False, want_near=False, ld32=OBS_LD):
    states = as_dev(states).reshape(-1, 5)
    B = states.shape[0]
    goals = as_dev(goals).reshape(B, -1)
    o64 = torch.empty(B, OBS_DIM, dtype=torch.float64, device=states.device) if want64 else None
    o32 = torch.empty(B, ld32, dtype=torch.float32, device=states.device) if want32 else None
    near = torch.empty(B, dtype=torch.uint8, device=states.device) if want_near else None
    check(lib().crlk_local_obs(env.h, _ptr(states), _ptr(goals), goals.shape[1], B, _ptr(o64),
                               _ptr(o32), ld32, _ptr(near), _stream()))
    return o64, o32, near


def gym_step(env, states, goals, actions, reward_param, step_dt=0.2, want64=True, want32=False,
             want_near=False, ld32=OBS_LD):
    """DifferentialDriveGym.step for B environments (differential_gym.py:136-178).  `states`
    (float64 CUDA [B, 5]) is advanced IN PLACE; actions None = info / reward / obs of the current
    states.  Returns (info [B, 8], reward [B], obs64, obs32, near)."""
    assert states.is_cuda and states.dtype == torch.float64 and states.is_contiguous()
    B = states.shape[0]
    goals = as_dev(goals).reshape(B, -1)
    actions = None if actions is None else as_dev(actions).reshape(B, 2)
    d = states.device
    info = torch.empty(B, 8, dtype=torch.float64, device=d)
    reward = torch.empty(B, dtype=torch.float64, device=d)
    o64 = torch.empty(B, OBS_DIM, dtype=torch.float64, device=d) if want64 else None
    o32 = torch.empty(B, ld32, dtype=torch.float32, device=d) if want32 else None
    near = torch.empty(B, dtype=torch.uint8, device=d) if want_near else None
    rp = (C.c_double * 8)(*[float(x) for x in np.asarray(reward_param, dtype=np.float64).reshape(8)])
    check(lib().crlk_gym_step(env.h, _ptr(states), _ptr(goals), goals.shape[1], _ptr(actions), float(step_dt),
                              rp, B, _ptr(info), _ptr(reward), _ptr(o64), _ptr(o32), ld32, _ptr(near),
                              _stream()))
    return info, reward, o64, o32, near


def dwa_traj_rows(env, dp):
    r = lib().crlk_dwa_traj_rows(env.h, C.byref(dp))
    if r < 0:
        check(r)
    return r


def dwa_control(env, dp, states, goals, want_traj=False):
    states = as_dev(states).reshape(-1, 5)
    B = states.shape[0]
    goals = as_dev(goals).reshape(B, -1)
    d = states.device
    opt_v = torch.empty(B, 2, dtype=torch.float64, device=d)
    idx = torch.empty(B, dtype=torch.int32, device=d)
    cost = torch.empty(B, dtype=torch.float64, device=d)
    nroll = torch.empty(B, dtype=torch.int32, device=d)
    tie = torch.empty(B, dtype=torch.uint8, device=d)
    traj = None
    if want_traj:
        traj = torch.empty(B, dwa_traj_rows(env, dp), 5, dtype=torch.float64, device=d)
    check(lib().crlk_dwa_control(env.h, C.byref(dp), _ptr(states), _ptr(goals), goals.shape[1], B,
                                 _ptr(opt_v), _ptr(idx), _ptr(cost), _ptr(traj), _ptr(nroll), _ptr(tie),
                                 _stream()))
    return dict(opt_v=opt_v, opt_idx=idx, opt_cost=cost, opt_traj=traj, n_rollouts=nroll, near_tie=tie)


def nearest(x64, y64, n_nodes, targets, x32=None, y32=None, coord_bound=0.0, max_nodes=None, states=None):
    """x64/y64 (and optional float32 copies): [Q, stride] SoA trees on the device.  With `states`
    ([Q, stride, 5] float64) and x64 = y64 = None the exact coordinates are read from columns 0 / 1
    of the state table (no separate float64 coordinate arrays)."""
    if states is not None:
        Q, stride, _ = states.shape
        d = states.device
        px, py, step = C.c_void_p(states.data_ptr()), C.c_void_p(states.data_ptr() + 8), 5
    else:
        Q, stride = x64.shape
        d = x64.device
        px, py, step = _ptr(x64), _ptr(y64), 1
    targets = as_dev(targets).reshape(Q, -1)
    n_nodes = n_nodes.to(device=d, dtype=torch.int32).contiguous()
    if max_nodes is None:
        max_nodes = stride
    idx = torch.empty(Q, dtype=torch.int32, device=d)
    dist = torch.empty(Q, dtype=torch.float64, device=d)
    tie = torch.empty(Q, dtype=torch.uint8, device=d)
    nbytes = lib().crlk_nearest_scratch_bytes(Q, int(max_nodes))
    scratch = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=d)
    check(lib().crlk_nearest(px, py, step, _ptr(x32), _ptr(y32), stride, _ptr(n_nodes),
                             int(max_nodes), float(coord_bound), _ptr(targets), targets.shape[1], Q,
                             _ptr(idx), _ptr(dist), _ptr(tie), _ptr(scratch), _stream()))
    return idx, dist, tie


def extend_dwa(env, dp, xp, from_states, targets, active=None, want_ctrl=False, out=None, stats=None):
    from_states = as_dev(from_states).reshape(-1, 5)
    Q = from_states.shape[0]
    targets = as_dev(targets).reshape(Q, 5)
    d = from_states.device
    per = xp.n_max_extend // xp.n_tree + 1
    if out is None:
        out = dict(path=torch.empty(Q, xp.n_max_extend, 5, dtype=torch.float64, device=d),
                   n_steps=torch.empty(Q, dtype=torch.int32, device=d),
                   status=torch.empty(Q, dtype=torch.int32, device=d),
                   node_step=torch.zeros(Q, per, dtype=torch.int32, device=d),
                   n_nodes=torch.empty(Q, dtype=torch.int32, device=d),
                   flags=torch.empty(Q, dtype=torch.uint8, device=d),
                   ctrl_idx=torch.full((Q, xp.n_max_extend), -1, dtype=torch.int32, device=d)
                   if want_ctrl else None)
    check(lib().crlk_extend_dwa(env.h, C.byref(dp), C.byref(xp), _ptr(from_states), _ptr(targets), Q,
                                _ptr(out["path"]), _ptr(out["n_steps"]), _ptr(out["status"]),
                                _ptr(out["node_step"]), _ptr(out["n_nodes"]), _ptr(out.get("ctrl_idx")),
                                _ptr(out["flags"]), _ptr(active), _ptr(stats), _stream()))
    return out


def gather_rows(src, idx, out=None):
    """src: [Q, N, C] float64 CUDA, idx: int32 [Q] -> [Q, C] (src[q, idx[q]])."""
    Q, N, Cc = src.shape
    if out is None:
        out = torch.empty(Q, Cc, dtype=torch.float64, device=src.device)
    check(lib().crlk_gather_rows(_ptr(src), N * Cc, Cc, _ptr(idx), Q, _ptr(out), _stream()))
    return out


# ------------------------------------------------------------------ networks

class PolicyHandle:
    """Owns a CrlkPolicy* (actor weights on the device)."""

    def __init__(self, weights, biases, low, high):
        require_cuda()
        n = len(weights)
        ws = [np.ascontiguousarray(w, dtype=np.float32) for w in weights]
        bs = [np.ascontiguousarray(b, dtype=np.float32) for b in biases]
        dims = [ws[0].shape[1]] + [w.shape[0] for w in ws]
        self.dims = dims
        dims_c = (C.c_int32 * (n + 1))(*dims)
        wp = (C.c_void_p * n)(*[w.ctypes.data for w in ws])
        bp = (C.c_void_p * n)(*[b.ctypes.data for b in bs])
        lo = np.ascontiguousarray(low, dtype=np.float32)
        hi = np.ascontiguousarray(high, dtype=np.float32)
        h = C.c_void_p()
        check(lib().crlk_policy_create(n, dims_c, wp, bp, lo.ctypes.data_as(C.c_void_p),
                                       hi.ctypes.data_as(C.c_void_p), C.byref(h)))
        self.h = h
        self.low, self.high = lo, hi

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().crlk_policy_destroy(self.h)
                self.h = None
        except Exception:
            pass


def pad_obs(obs, ld=OBS_LD):
    """[B, 733] (numpy or tensor, any float dtype) -> float32 CUDA [B, ld], zero padded."""
    o = as_dev(obs, torch.float32)
    o = o.reshape(-1, o.shape[-1])
    if o.shape[1] == ld:
        return o
    out = torch.zeros(o.shape[0], ld, dtype=torch.float32, device=o.device)
    out[:, :o.shape[1]] = o
    return out


def policy_forward(pol, obs32, noise=None, eps=0.0):
    """obs32: float32 CUDA [B, ld>=736, zero padded]; noise: float32 [B, 2] or None."""
    B, ld = obs32.shape
    act = torch.empty(B, 2, dtype=torch.float32, device=obs32.device)
    nbytes = lib().crlk_policy_workspace_bytes(pol.h, B)
    ws = torch.empty(int(nbytes), dtype=torch.uint8, device=obs32.device)
    if noise is not None:
        noise = as_dev(noise, torch.float32).reshape(B, 2)
    check(lib().crlk_policy_forward(pol.h, _ptr(obs32), ld, _ptr(noise), float(eps), B, _ptr(act),
                                    _ptr(ws), _stream()))
    return act


class EstimatorHandle:
    """Owns a CrlkEstimator* (estimator + classifier CNN weights)."""
    ORDER = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias",
             "conv4.weight", "conv4.bias", "linear.weight", "linear.bias"]

    def __init__(self, est_sd, cls_sd):
        require_cuda()
        def pack(sd):
            arrs = [np.ascontiguousarray(np.asarray(sd[k], dtype=np.float32)) for k in self.ORDER]
            return arrs, (C.c_void_p * 10)(*[a.ctypes.data for a in arrs])
        ea, ep = pack(est_sd)
        ca, cp = pack(cls_sd)
        h = C.c_void_p()
        check(lib().crlk_estimator_create(ep, cp, C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().crlk_estimator_destroy(self.h)
                self.h = None
        except Exception:
            pass


def estimator_forward(est, obs32):
    B, ld = obs32.shape
    e = torch.empty(B, dtype=torch.float32, device=obs32.device)
    p = torch.empty(B, dtype=torch.float32, device=obs32.device)
    check(lib().crlk_estimator_forward(est.h, _ptr(obs32), ld, B, _ptr(e), _ptr(p), _stream()))
    return e, p


def estimator_choose_parent(env, est, states, n_nodes, targets, max_nodes=None, first_node=None):
    """states: float64 CUDA [Q, stride, 5]; n_nodes int32 [Q]; targets [Q, 5] -> (idx i32[Q], cost f64[Q]).
    first_node (int32 [Q], optional): candidates are the nodes [first_node[q], n_nodes[q])."""
    Q, stride, _ = states.shape
    targets = as_dev(targets).reshape(Q, 5)
    n_nodes = n_nodes.to(device=states.device, dtype=torch.int32).contiguous()
    max_nodes = int(max_nodes or stride)
    idx = torch.empty(Q, dtype=torch.int32, device=states.device)
    cost = torch.empty(Q, dtype=torch.float64, device=states.device)
    nbytes = lib().crlk_estimator_cp_workspace_bytes(Q, max_nodes)
    ws = torch.empty(int(nbytes), dtype=torch.uint8, device=states.device)
    check(lib().crlk_estimator_choose_parent(env.h, est.h, _ptr(states), stride, _ptr(n_nodes), _ptr(first_node),
                                             max_nodes, _ptr(targets), Q, _ptr(idx), _ptr(cost), _ptr(ws),
                                             _stream()))
    return idx, cost


def extend_rl_outputs(xp, Q, d, want_actions=False):
    per = xp.n_max_extend // xp.n_tree + 1
    return dict(path=torch.empty(Q, xp.n_max_extend, 5, dtype=torch.float64, device=d),
                n_steps=torch.empty(Q, dtype=torch.int32, device=d),
                status=torch.empty(Q, dtype=torch.int32, device=d),
                node_step=torch.zeros(Q, per, dtype=torch.int32, device=d),
                n_nodes=torch.empty(Q, dtype=torch.int32, device=d),
                flags=torch.empty(Q, dtype=torch.uint8, device=d),
                actions=torch.zeros(Q, xp.n_max_extend, 2, dtype=torch.float32, device=d)
                if want_actions else None)


def extend_rl(env, pol, xp, from_states, targets, noise=None, eps=0.0, active=None, want_actions=False,
              noise_cursor=None, out=None, workspace=None):
    """RRT_RL.steer (rrt_rl.py:22-63) for Q (from, target) pairs.  out / workspace: reuse the output
    tensors / the scratch buffer of an earlier call (same Q or larger workspace)."""
    from_states = as_dev(from_states).reshape(-1, 5)
    Q = from_states.shape[0]
    targets = as_dev(targets).reshape(Q, 5)
    d = from_states.device
    if out is None:
        out = extend_rl_outputs(xp, Q, d, want_actions)
    noise_rows = 0
    if noise is not None:
        noise = as_dev(noise, torch.float32).reshape(Q, -1, 2)
        noise_rows = noise.shape[1]
    nbytes = int(lib().crlk_extend_rl_workspace_bytes(pol.h, Q)) + 256
    ws = workspace if workspace is not None and workspace.numel() >= nbytes else out.get("_ws")
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=d)
    off = (-ws.data_ptr()) % 256
    check(lib().crlk_extend_rl(env.h, pol.h, C.byref(xp), _ptr(from_states), _ptr(targets), Q,
                               _ptr(noise), noise_rows, _ptr(noise_cursor), float(eps), _ptr(out["path"]),
                               _ptr(out["n_steps"]),
                               _ptr(out["status"]), _ptr(out["node_step"]), _ptr(out["n_nodes"]),
                               _ptr(out.get("actions")), _ptr(out["flags"]), _ptr(active),
                               C.c_void_p(ws.data_ptr() + off), _stream()))
    out["_ws"] = ws
    return out
