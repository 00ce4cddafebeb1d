"""CPU-only checks of the drop-in boundary: libcrlk.so loads without a GPU,
exports every symbol include/crlk.h declares, fails loudly (no CPU fallback),
and the product package never touches oracle/."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "crlk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crlk_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from crl_kino_b200._lib import PROTOTYPES, SO, lib
    names = _declared()
    assert len(names) >= 25
    raw = ctypes.CDLL(SO)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in crlk.h but not exported by libcrlk.so"
        assert n in PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(PROTOTYPES) == names
    assert lib().crlk_abi_version() == 2


def test_struct_layouts_match_header():
    from crl_kino_b200._lib import DwaParams, EnvParams, ExtendParams, PlanParams
    assert ctypes.sizeof(EnvParams) == 6 * 8 + 4 * 8 + 4 * 4 + 8
    assert ctypes.sizeof(PlanParams) == 8 + 4 * 8 + 8 + 10 * 8 + 8
    assert ctypes.sizeof(DwaParams) == 7 * 8 + 2 * 4
    assert ctypes.sizeof(ExtendParams) == 4 + 4 + 8 + 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from crl_kino_b200._lib import CrlkError, EnvParams, lib
    l = lib()
    assert l.crlk_device_count() == 0
    p = EnvParams(1.0, -0.1, np.pi, 1.0, np.pi, 0.1, (ctypes.c_double * 4)(-20, 20, -20, 20),
                  (ctypes.c_float * 4)(-0.7, -0.5, 0.7, 0.5), 0.0)
    h = ctypes.c_void_p()
    rc = l.crlk_env_create(ctypes.byref(p), None, 0, ctypes.byref(h))
    assert rc == -3 and b"no CPU fallback" in l.crlk_last_error()
    from crl_kino_b200.env import DifferentialDriveEnv
    env = DifferentialDriveEnv()
    with pytest.raises(CrlkError):
        env.valid_state_check(np.zeros(5))
    from crl_kino_b200.planner.rrt import RRT
    with pytest.raises(CrlkError):
        RRT(env).set_start_and_goal(np.zeros(5), np.ones(5))


def test_invalid_arguments_are_reported():
    from crl_kino_b200._lib import lib
    l = lib()
    assert l.crlk_env_create(None, None, 0, None) == -1
    assert b"invalid argument" in l.crlk_last_error()
    assert l.crlk_nearest_scratch_bytes(0, 10) == 0
    assert l.crlk_nearest_scratch_bytes(4, 100000) > 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "crl_kino_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"


def test_host_side_containers_and_workloads():
    from crl_kino_b200.env import DifferentialDriveEnv, DifferentialDriveGym
    from crl_kino_b200.utils.rigid import RectObs, rects_of
    from crl_kino_b200.workloads import dense_map, sample_stream
    from conftest import golden
    fx = golden("fixtures")
    env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi)
    assert np.array_equal(env.get_bounds()["state_bounds"], fx["state_bounds"])
    assert np.array_equal(env.rigid_robot.rect.reshape(4), fx["robot_rect"])
    gym = DifferentialDriveGym(env)
    assert np.array_equal(gym.sample_positions, fx["sample_positions"])
    assert np.array_equal(gym.action_space.low, fx["action_low"]) and gym.action_space.low.dtype == np.float32
    obs = [RectObs(r.reshape(2, 2)) for r in fx["map_test_env1"]]
    env.set_obs(obs)
    assert np.array_equal(rects_of(env.obs_list), fx["map_test_env1"]) and env.rects.dtype == np.float32
    with pytest.raises(AssertionError):
        env.add_obs(object())
    m = dense_map()
    assert m.shape == (10004, 4) and m.dtype == np.float32 and np.array_equal(dense_map(), m)
    s = sample_stream(fx["state_bounds"], np.array([10, 10, 0, 0, 0.0]), 4000, seed=3)
    frac_goal = np.mean(np.all(s == np.array([10, 10, 0, 0, 0.0]), axis=1))
    assert 0.03 < frac_goal < 0.09          # (5 + 1) / 101 goal bias (rrt.py:147)


def test_reference_pickle_loader():
    """utils.rigid.load_obstacles reads pickles of crl_kino.utils.rigid.RectObs."""
    import pickle
    import sys
    import types
    from crl_kino_b200.utils import rigid
    mod = types.ModuleType("crl_kino.utils.rigid")
    exec("import numpy as np\n"
         "class RectObs:\n"
         "    def __init__(self, rect):\n"
         "        self.rect = np.array(rect, dtype=np.float32)\n"
         "        self.color = 'black'\n", mod.__dict__)
    sys.modules["crl_kino"] = types.ModuleType("crl_kino")
    sys.modules["crl_kino.utils"] = types.ModuleType("crl_kino.utils")
    sys.modules["crl_kino.utils.rigid"] = mod
    try:
        blob = pickle.dumps([mod.RectObs([[0, 0], [1, 2]]), mod.RectObs([[-3, -3], [-1, -2]])])
    finally:
        for k in ("crl_kino", "crl_kino.utils", "crl_kino.utils.rigid"):
            sys.modules.pop(k, None)
    obs = rigid.load_obstacles(blob)
    assert isinstance(obs[0], rigid.RectObs)
    assert np.array_equal(rigid.rects_of(obs), np.array([[0, 0, 1, 2], [-3, -3, -1, -2]], dtype=np.float32))
