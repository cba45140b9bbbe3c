import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


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
