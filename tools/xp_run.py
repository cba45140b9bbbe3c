#!/usr/bin/env python3
"""Time the bench workload with each A/B library under tools/xp/ (BRAWL_CUDA_LIB): python tools/xp_run.py [layout] [names...]
One subprocess per library; prints swaps/s (best of 3 runs of 64 sweeps, device-resident)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time, os
sys.path.insert(0, %r)
import numpy as np, brawl_b200 as bw, bench
n = bench.N_CELLS
dev = bw.Device("bcc", n, n, n, 4, 4, bench.load_V())
dev.metropolis_set_layout(int(sys.argv[1]))
dev.set_config(bench.synthetic_config(n, 4, 0))
beta = 1.0 / (bench.T_KELVIN * bw.K_B_IN_RY)
dev.metropolis_run(beta, 32 * dev.n_atoms)
best = 0.0
for _ in range(4):
    t0 = time.perf_counter()
    att, acc, dE = dev.metropolis_run(beta, 128 * dev.n_atoms)
    best = max(best, att[0] / (time.perf_counter() - t0))
print("%%-24s %%.4e swaps/s  acceptance %%.4f" %% (os.path.basename(os.environ.get("BRAWL_CUDA_LIB", "default")), best, acc[0] / att[0]), flush=True)
''' % ROOT
layout = sys.argv[1] if len(sys.argv) > 1 else "0"
names = sys.argv[2:] or sorted(f[3:-3] for f in os.listdir(os.path.join(ROOT, "tools", "xp")) if f.endswith(".so"))
for rep in range(2):
    for nm in ["default"] + names:
        env = dict(os.environ)
        if nm != "default":
            env["BRAWL_CUDA_LIB"] = os.path.join(ROOT, "tools", "xp", "lib%s.so" % nm)
        subprocess.run([sys.executable, "-c", CHILD, layout], env=env)
