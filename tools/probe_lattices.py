#!/usr/bin/env python3
"""Throughput of the production Metropolis kernels on 128^3-cell lattices other than the headline one (fcc, 6 shells)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brawl_b200 as bw  # noqa: E402

gold = np.load(os.path.join(ROOT, "tests", "golden", "brawl_golden.npz"))
CASES = [("bcc", 128, 4, 4, "ex_AlTiCrMo_V"), ("bcc", 128, 4, 6, "t02_V"), ("fcc", 128, 5, 4, "ex_AlCrFeCoNi_V"), ("fcc", 128, 2, 4, "ex_FeNi_V"),
         ("fcc", 128, 5, 6, "t01_V")]
for lattice, n, S, shells, key in CASES:
    V = np.ascontiguousarray(gold[key][: S * S * shells])
    dev = bw.Device(lattice, n, n, n, S, shells, V)
    N = dev.n_atoms
    rng = np.random.default_rng(1)
    g = np.zeros((2 * n,) * 3, dtype=np.int8)
    z, y, x = np.meshgrid(np.arange(2 * n, dtype=np.int16), np.arange(2 * n, dtype=np.int16), np.arange(2 * n, dtype=np.int16), indexing="ij", sparse=True)
    mask = ((x & 1) == (z & 1)) & ((y & 1) == (z & 1)) if lattice == "bcc" else ((x + y + z) & 1) == 0
    g[mask] = rng.integers(1, S + 1, size=int(mask.sum()))
    dev.set_config(g)
    beta = 1.0 / (1000.0 * bw.K_B_IN_RY)
    plan = dev.metropolis_plan()
    dev.metropolis_run(beta, 4 * N)
    best = 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        att, acc, dE = dev.metropolis_run(beta, 16 * N)
        best = max(best, att[0] / (time.perf_counter() - t0))
    print(json.dumps({"lattice": lattice, "n": n, "S": S, "shells": shells, "swaps_per_s": best, "acceptance": float(acc[0]) / float(att[0]), "plan": plan}), flush=True)
    dev.close()
