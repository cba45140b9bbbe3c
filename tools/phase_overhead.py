#!/usr/bin/env python3
"""Fixed cost of a phase (= one launch: box load + store, table set-up, counter reduction) of the production kernel on the
bench workload, from launches with different numbers of steps: t(steps) = overhead + steps * t_step."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brawl_b200 as bw   # noqa: E402
import bench              # noqa: E402

n = bench.N_CELLS
beta = 1.0 / (bench.T_KELVIN * bw.K_B_IN_RY)
rows = []
for steps in (52, 104, 208, 312):
    dev = bw.Device("bcc", n, n, n, 4, 4, bench.load_V())
    dev.metropolis_tune((0, 0, 0), steps)
    dev.set_config(bench.synthetic_config(n, 4, 0))
    plan = dev.metropolis_plan()
    dev.metropolis_run(beta, 16 * dev.n_atoms)
    best = 1e9
    for _ in range(4):
        t0 = time.perf_counter()
        att, acc, dE = dev.metropolis_run(beta, 64 * dev.n_atoms)
        dt = time.perf_counter() - t0
        best = min(best, dt / dev.metropolis_last_launches())
    rows.append((plan["steps_per_phase"], best * 1e6, att[0] / dt))
    print("steps/phase %4d: %.1f us per launch, %.3e swaps/s" % (plan["steps_per_phase"], best * 1e6, att[0] / dt), flush=True)
    dev.close()
(s0, t0, _), (s1, t1, _) = rows[0], rows[-1]
t_step = (t1 - t0) / (s1 - s0)
for s, t, _ in rows:
    print("steps %4d: overhead %.1f us = %.1f %% of the launch" % (s, t - s * t_step, 100 * (t - s * t_step) / t))
