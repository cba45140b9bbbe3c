#!/usr/bin/env python3
"""A few phases of the production Metropolis kernel on the bench workload (128^3 bcc AlTiCrMo, 1000 K) for ncu:
    python tools/prof_run.py [layout=0] [sweeps=24] [steps_per_phase=0] [workload=chain|bcc6|fcc4|feni|fcc6]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brawl_b200 as bw        # noqa: E402
import bench                   # noqa: E402

layout = int(sys.argv[1]) if len(sys.argv) > 1 else 0
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
workload = sys.argv[4] if len(sys.argv) > 4 else "chain"
n = bench.N_CELLS
if workload == "chain":
    dev = bw.Device("bcc", n, n, n, 4, 4, bench.load_V())
    g0 = bench.synthetic_config(n, 4, 0)
else:
    lattice, S0, S, shells, key, _ = bench.OTHER_LATTICES[workload]
    gold = np.load(os.path.join(ROOT, "tests", "golden", "brawl_golden.npz"))
    V = np.ascontiguousarray(gold[key][: S0 * S0 * shells].reshape(shells, S0, S0)[:, :S, :S]).ravel()
    dev = bw.Device(lattice, n, n, n, S, shells, V)
    g0 = bench.synthetic_config(n, S, 0, lattice)
dev.metropolis_set_layout(layout)
if steps:
    dev.metropolis_tune((0, 0, 0), steps)
dev.set_config(g0)
att, acc, dE = dev.metropolis_run(1.0 / (bench.T_KELVIN * bw.K_B_IN_RY), sweeps * dev.n_atoms)
print(dev.metropolis_plan(), int(att[0]), int(acc[0]), dev.metropolis_last_launches())
