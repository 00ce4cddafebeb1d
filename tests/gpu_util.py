"""Shared helpers of the GPU parity tests."""
import numpy as np

from conftest import golden
from oracle import oracle as orc


def make_envs(rects):
    """(product env, oracle env) over the same float32 obstacle array."""
    from crl_kino_b200.env import DifferentialDriveEnv
    env = DifferentialDriveEnv(1.0, -0.1, np.pi, 1.0, np.pi)
    env.set_obs(np.asarray(rects, dtype=np.float32))
    o = orc.Env()
    o.set_obs(rects)
    return env, o


def fixture_map(name):
    return golden("fixtures")["map_" + name]


def np_(t):
    return t.detach().cpu().numpy()


def assert_flags(got, ref, near, what):
    """Flags must be bit-exact except where the kernel reported a near-boundary decision."""
    got, ref, near = np.asarray(got).astype(bool), np.asarray(ref).astype(bool), np.asarray(near).astype(bool)
    bad = (got != ref) & ~near
    assert not bad.any(), f"{what}: {int(bad.sum())} unexplained flag mismatches at {np.where(bad)[0][:8]}"
    return int(((got != ref) & near).sum())
