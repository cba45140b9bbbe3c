"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/brawl_cuda.h declares; with no GPU it fails loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "brawl_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(brawl_cuda_\w+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    from brawl_b200 import build
    build.build_library()
    import brawl_b200
    return brawl_b200.load()


def test_header_symbols_are_exported(lib):
    import brawl_b200
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "libbrawl_cuda.so does not export %s" % s
    assert sorted(brawl_b200.EXPORTS) == syms, "ctypes binding and header disagree"


def test_version(lib):
    assert lib.brawl_cuda_version() == 1


def test_product_does_not_reference_oracle():
    """The product path must never import/link the CPU oracle."""
    for base, _, files in os.walk(os.path.join(ROOT, "brawl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                for line in open(os.path.join(base, f)).read().splitlines():
                    code = line.split("#")[0] if f.endswith(".py") else line.split("//")[0]
                    assert "liboracle" not in code and "import oracle" not in code and "from oracle" not in code, (f, line)


def test_fails_loudly_without_gpu(lib):
    import brawl_b200
    n = C.c_int(0)
    if lib.brawl_cuda_device_count(C.byref(n)) == 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(brawl_b200.BrawlCudaError):
        brawl_b200.Device("bcc", 4, 4, 4, 4, 4, np.zeros(64))


def test_argument_validation_precedes_device_use(lib):
    import brawl_b200
    with pytest.raises(brawl_b200.BrawlCudaError, match="Unsupported number of shells"):
        brawl_b200.Device("fcc", 4, 4, 4, 2, 7, np.zeros(2 * 2 * 7))
    with pytest.raises(brawl_b200.BrawlCudaError, match="Unsupported number of shells"):
        brawl_b200.Device("simple_cubic", 4, 4, 4, 2, 3, np.zeros(2 * 2 * 3))
