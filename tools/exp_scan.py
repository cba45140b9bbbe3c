#!/usr/bin/env python3
"""GPU scan: throughput (128^3) and sampling efficiency (relaxation from a random start at 1000 K against the oracle's
sequential sampler) of the word kernels as a function of the epoch length K and the number of steps per phase.

efficiency = (sweeps the sequential sampler needs to reach E*) / (sweeps this kernel needs), E* = the sequential
sampler's energy after `REF_SWEEPS` sweeps; physical rate = attempts/s x efficiency."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brawl_b200 as bw                      # noqa: E402
from oracle import oracle as orc             # noqa: E402  (checker only)
from tools.exp_epoch import rand_config, V4  # noqa: E402

T = 1000.0
BETA = 1.0 / (T * bw.K_B_IN_RY)
REF_MARKS = (8, 16, 32)


def out(**kw):
    print(json.dumps(kw), flush=True)


def oracle_curve(n, chains=4, sweeps=32):
    sysm = orc.System("bcc", n, n, n, 4, 4, V4)
    N = sysm.n_atoms
    curves = []
    for c in range(chains):
        g = rand_config(n, 4, 100 + c)
        mt = orc.MT(seed=900 + c)
        cv = [(0.0, sysm.total_energy(g) / N)]
        for m in range(1, sweeps + 1):
            sysm.metropolis_trials(g, mt, BETA, N)
            cv.append((float(m), sysm.total_energy(g) / N))
        curves.append(cv)
    a = np.array(curves)                      # [chains][marks][2]
    return a[0, :, 0], a[:, :, 1].mean(axis=0), a[:, :, 1].std(axis=0, ddof=1) / np.sqrt(chains)


def gpu_curve(n, layout, steps, R=8, max_sweeps=120.0):
    dev = bw.Device("bcc", n, n, n, 4, 4, V4, n_replicas=R)
    dev.metropolis_set_layout(layout)
    if steps:
        dev.metropolis_tune((0, 0, 0), steps)
    N = dev.n_atoms
    dev.set_config(np.stack([rand_config(n, 4, 100 + r) for r in range(R)]))
    xs, es = [0.0], [float((dev.total_energy(0, R, exact_order=False) / N).mean())]
    tot = 0
    while xs[-1] < max_sweeps:
        att, acc, dE = dev.metropolis_run(BETA, 1, seed=7)          # one phase
        tot += int(att[0])
        xs.append(tot / N)
        es.append(float((dev.total_energy(0, R, exact_order=False) / N).mean()))
    return np.array(xs), np.array(es)


def sweeps_to_reach(xs, es, target):
    """first crossing of `target` (energies decrease), linear interpolation; None if never reached"""
    for i in range(1, len(xs)):
        if es[i] <= target:
            f = (es[i - 1] - target) / (es[i - 1] - es[i])
            return float(xs[i - 1] + f * (xs[i] - xs[i - 1]))
    return None


def throughput(layout, steps):
    n = 128
    g = rand_config(n, 4, 1)
    N = 2 * n ** 3
    dev = bw.Device("bcc", n, n, n, 4, 4, V4)
    dev.metropolis_set_layout(layout)
    if steps:
        dev.metropolis_tune((0, 0, 0), steps)
    plan = dev.metropolis_plan()
    dev.set_config(g)
    dev.metropolis_run(BETA, 16 * N)
    best = 0.0
    for rep in range(3):
        t0 = time.perf_counter()
        att, acc, dE = dev.metropolis_run(BETA, 48 * N)
        dt = time.perf_counter() - t0
        best = max(best, att[0] / dt)
    return best, plan["steps_per_phase"], dev.metropolis_last_launches()


if __name__ == "__main__":
    n = 32
    ox, oe, ose = oracle_curve(n)
    out(test="oracle_curve", sweeps=ox.tolist(), e=oe.tolist(), se=ose.tolist())
    targets = {m: float(oe[m]) for m in REF_MARKS}
    configs = [("word_split(r01)", 3, s) for s in (0, 60)]
    for name, lay in (("epoch2", 5), ("epoch4", 0), ("epoch8", 4)):
        configs += [(name, lay, s) for s in ((0, 104, 64, 32) if len(sys.argv) < 2 else (0, 104))]
    for name, lay, steps in configs:
        rate, sp, launches = throughput(lay, steps)
        xs, es = gpu_curve(n, lay, steps)
        eff = {}
        for m, tgt in targets.items():
            s = sweeps_to_reach(xs, es, tgt)
            eff[m] = None if s is None else m / s
        out(test="scan", layout=name, steps_per_phase=sp, swaps_per_s=rate, efficiency=eff,
            physical_rate={m: (None if e is None else rate * e) for m, e in eff.items()},
            curve=[(round(float(a), 2), float(b)) for a, b in zip(xs[:40], es[:40])])
