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


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the oracle port on the host cores; no GPU needed): one JSON line with the keys the driver
    reads, the same workload string as the CUDA arm, impl == "reference", e2e == value with zero copies."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-trials", "2000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"]["workload"] == bench.workload_name(bench.N_CELLS) and d["metric"] == bench.METRIC
