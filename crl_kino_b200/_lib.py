"""ctypes binding of include/crlk.h.  There is no CPU fallback: if libcrlk.so is
missing or no CUDA device is present every compute call raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libcrlk.so")


class CrlkError(RuntimeError):
    pass


class EnvParams(C.Structure):
    _fields_ = [("max_v", C.c_double), ("min_v", C.c_double), ("max_w", C.c_double),
                ("max_acc_v", C.c_double), ("max_acc_w", C.c_double), ("base_dt", C.c_double),
                ("env_bounds", C.c_double * 4), ("robot_rect", C.c_float * 4), ("robot_radius", C.c_double)]


class DwaParams(C.Structure):
    _fields_ = [("v_res", C.c_double), ("w_res", C.c_double), ("dt", C.c_double),
                ("predict_time", C.c_double), ("to_goal_cost_gain", C.c_double),
                ("ob_gain", C.c_double), ("speed_cost_gain", C.c_double),
                ("skip", C.c_int32), ("use_clearance", C.c_int32)]


class ForestView(C.Structure):
    _fields_ = [("x64", C.c_void_p), ("y64", C.c_void_p), ("x32", C.c_void_p), ("y32", C.c_void_p),
                ("states", C.c_void_p), ("parent", C.c_void_p), ("n_nodes", C.c_void_p),
                ("stride", C.c_int64), ("pool", C.c_void_p), ("seg_off", C.c_void_p),
                ("seg_len", C.c_void_p), ("seg_flag", C.c_void_p), ("pool_used", C.c_void_p),
                ("pool_cap", C.c_int64)]


class ExtendView(C.Structure):
    _fields_ = [("path", C.c_void_p), ("n_steps", C.c_void_p), ("status", C.c_void_p),
                ("node_step", C.c_void_p), ("n_nodes", C.c_void_p)]


class PlanParams(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("d_min", C.c_double), ("d_max", C.c_double),
                ("goal_radius", C.c_double), ("retry_radius", C.c_double), ("seed", C.c_uint64),
                ("state_lo", C.c_double * 5), ("state_hi", C.c_double * 5), ("goal_sample_rate", C.c_int32)]


class ExtendParams(C.Structure):
    _fields_ = [("n_max_extend", C.c_int32), ("n_tree", C.c_int32), ("step_dt", C.c_double),
                ("reach_radius", C.c_double)]


P = C.c_void_p
I32 = C.c_int32
I64 = C.c_int64
F64 = C.c_double
F32 = C.c_float

# name -> (restype, argtypes): one entry per declaration in include/crlk.h
PROTOTYPES = {
    "crlk_last_error": (C.c_char_p, []),
    "crlk_abi_version": (C.c_int, []),
    "crlk_device_count": (C.c_int, []),
    "crlk_env_create": (C.c_int, [P, P, I32, P]),
    "crlk_env_destroy": (C.c_int, [P]),
    "crlk_env_set_obs": (C.c_int, [P, P, I32]),
    "crlk_env_num_obs": (C.c_int, [P]),
    "crlk_valid_state": (C.c_int, [P, P, I32, P, P, P, P]),
    "crlk_valid_points": (C.c_int, [P, P, I32, P, P]),
    "crlk_motion_velocity": (C.c_int, [P, P, P, F64, I32, P, P]),
    "crlk_motion": (C.c_int, [P, P, P, F64, I32, P, P]),
    "crlk_dynamic_window": (C.c_int, [P, P, F64, I32, P, P]),
    "crlk_local_obs": (C.c_int, [P, P, P, I32, I32, P, P, I32, P, P]),
    "crlk_dwa_control": (C.c_int, [P, P, P, P, I32, I32, P, P, P, P, P, P, P]),
    "crlk_dwa_traj_rows": (C.c_int, [P, P]),
    "crlk_dwa_max_rollouts": (C.c_int, [P, P]),
    "crlk_nearest": (C.c_int, [P, P, I32, P, P, I64, P, I32, F64, P, I32, I32, P, P, P, P, P]),
    "crlk_nearest_scratch_bytes": (I64, [I32, I32]),
    "crlk_extend_dwa": (C.c_int, [P, P, P, P, P, I32, P, P, P, P, P, P, P, P, P, P]),
    "crlk_gather_rows": (C.c_int, [P, I64, I32, P, I32, P, P]),
    "crlk_estimator_choose_parent": (C.c_int, [P, P, P, I64, P, P, I32, P, I32, P, P, P, P]),
    "crlk_rrt_retry_choice": (C.c_int, [P, P, P, P, P, I32, F64, I32, P, P, P, P, P, P]),
    "crlk_estimator_cp_workspace_bytes": (I64, [I32, I32]),
    "crlk_rrt_select": (C.c_int, [P, P, I32, I32, F64, F64, P, P, P, P]),
    "crlk_rrt_append": (C.c_int, [P, P, P, P, P, I32, I32, I32, I32, I32, I32, F64, F64, P, P, P, P, P, P, P, P]),
    "crlk_debug_atan2": (C.c_int, [P, P, I32, P, P]),
    "crlk_debug_div": (C.c_int, [P, P, I32, P, P, P]),
    "crlk_debug_sincos": (C.c_int, [P, I32, P, P, P]),
    "crlk_gym_step": (C.c_int, [P, P, P, I32, P, F64, P, I32, P, P, P, P, I32, P, P]),
    "crlk_plan_dwa": (C.c_int, [P, P, P, P, P, P, P, I32, I32, P, P, P, P, P, P, P, P]),
    "crlk_sample_stream": (C.c_int, [P, P, I32, I32, P, P]),
    "crlk_policy_create": (C.c_int, [I32, P, P, P, P, P, P]),
    "crlk_policy_destroy": (C.c_int, [P]),
    "crlk_policy_forward": (C.c_int, [P, P, I32, P, F32, I32, P, P, P]),
    "crlk_policy_workspace_bytes": (I64, [P, I32]),
    "crlk_extend_rl": (C.c_int, [P, P, P, P, P, I32, P, I32, P, F32, P, P, P, P, P, P, P, P, P, P]),
    "crlk_extend_rl_workspace_bytes": (I64, [P, I32]),
    "crlk_extend_rl_launches": (I32, [P, P, I32, I32]),
    "crlk_estimator_create": (C.c_int, [P, P, P]),
    "crlk_estimator_destroy": (C.c_int, [P]),
    "crlk_estimator_forward": (C.c_int, [P, P, I32, I32, P, P, P]),
}

_lib = None


def lib():
    """Load libcrlk.so (built by crl_kino_b200/build.py).  Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise CrlkError(f"{SO} is missing: run `python -m crl_kino_b200.build` "
                            "(there is no CPU fallback)")
        l = C.CDLL(SO)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise CrlkError(f"libcrlk error {rc}: {lib().crlk_last_error().decode()}")


def require_cuda():
    import torch
    if not torch.cuda.is_available() or lib().crlk_device_count() <= 0:
        raise CrlkError("crl_kino_b200 needs a CUDA device: there is no CPU fallback")
