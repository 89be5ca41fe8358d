import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "reference_outputs.npz")))


@pytest.fixture(scope="session")
def golden_selftest():
    return dict(np.load(os.path.join(GOLDEN, "reference_selftest_250.npz")))


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build libhgrid.so once per session if it is missing or stale (nvcc works without a GPU)."""
    import __graft_entry__ as g
    g.build()
    yield


def cuts_list(arr):
    return [((int(a), int(b)), (int(c), int(d))) for a, b, c, d in arr]


def rel_err(a, b):
    """max abs error divided by the value range of the reference result"""
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.ptp(b), 1e-300))
