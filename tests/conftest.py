import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


DATA = os.path.join(ROOT, "data")        # reference maps / start-goal sets / estimator weights (inputs, not expectations)


def golden(name):
    d = DATA if name in ("fixtures", "estimator_weights") else GOLDEN
    return np.load(os.path.join(d, name + ".npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def fixtures():
    return golden("fixtures")
