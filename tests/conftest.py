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


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def kat():
    return load_golden("kat_small.npz")


@pytest.fixture(scope="session")
def trace():
    return load_golden("train_trace.npz")


@pytest.fixture(scope="session")
def traj15():
    return load_golden("traj_d15.npz")


@pytest.fixture(scope="session")
def fwd47():
    return load_golden("forward_d47.npz")


@pytest.fixture(scope="session")
def evalm():
    return load_golden("eval_metrics.npz")
