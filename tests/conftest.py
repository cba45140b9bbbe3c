import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_present():
    """True if the CUDA driver reports at least one device (libcuda through ctypes: no torch import at collection time)."""
    import ctypes
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        return cu.cuInit(0) == 0 and cu.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    """Without a GPU the `gpu` tests are skipped (not failed with 'CUDA error 35'), so CPU-side regressions stay visible."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: run on the B200 box with -m gpu")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    """Reference golden vectors converted by tests/golden/make_golden.py."""
    return np.load(os.path.join(ROOT, "tests", "golden", "brawl_golden.npz"))


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle -- checker only; the product never imports it."""
    from oracle import oracle
    oracle.lib()
    return oracle
