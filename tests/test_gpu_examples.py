"""The example scripts (counterparts of the reference's examples/test.py and benchmark.py) run end
to end on the GPU and write files in the reference's formats."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, cwd=ROOT):
    r = subprocess.run([sys.executable] + args, cwd=cwd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_example_cases(tmp_path):
    out = _run(["examples/test.py", "--case", "rrt", "--max-iter", "25", "--out", str(tmp_path)])
    assert "rrt: reached=" in out
    from crl_kino_b200.utils import io as cio
    nodes = cio.load_tree(str(tmp_path / "rrt_tree.pkl"))
    path = cio.load_path(str(tmp_path / "rrt_path.pkl"))
    assert len(nodes) >= 2 and path.shape[1] == 5 and np.array_equal(path[0], [-5, -15, 0, 0, 0])
    assert "gym:" in _run(["examples/test.py", "--case", "gym"])
    assert "queries/s" in _run(["examples/test.py", "--case", "batch", "--queries", "32", "--max-iter", "20"])
    assert "rl_rrt_estimator: reached=" in _run(["examples/test.py", "--case", "rl_rrt_estimator", "--max-iter", "3"])


def test_example_benchmark(tmp_path):
    out = _run(["examples/benchmark.py", "--planner", "rrt", "--max-iter", "40", "--out", str(tmp_path)])
    assert "test_env1:" in out and "test_env2:" in out
    for name in ("env1", "env2"):
        data = pickle.load(open(tmp_path / f"planning_results_rrt_test_{name}.pkl", "rb"))
        assert set(data) == {"runtime", "path_len"} and len(data["runtime"]) == 12
        assert len(data["path_len"]) == int((data["runtime"] >= 0).sum())
