"""Build libbrawl_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbrawl_cuda.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inc")))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(os.path.dirname(HERE), "include", "brawl_cuda.h")]
    return any(os.path.getmtime(s) > t for s in deps)


def build_library(force=False, verbose=False):
    """Compile brawl_b200/csrc/brawl_cuda.cu (two translation units) -> brawl_b200/libbrawl_cuda.so."""
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # two translation units compiled side by side (no -rdc: kernels cross the boundary as host function pointers), then one
    # link.  NOT -split-compile: it changed the register allocation of unrelated kernels from build to build (measured).
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    units = ["brawl_cuda.cu", "byte_epoch_kernels.cu"]
    objs = [os.path.join(CSRC, u[:-3] + ".o") for u in units]
    procs = [subprocess.Popen([nvcc] + flags + ["-c", "-o", o, os.path.join(CSRC, u)], cwd=CSRC) for u, o in zip(units, objs)]
    rc = [p.wait() for p in procs]
    if any(rc):
        raise subprocess.CalledProcessError(max(rc), "nvcc -c " + " ".join(units))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs, cwd=CSRC)
    for o in objs:
        os.remove(o)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
