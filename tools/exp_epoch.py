#!/usr/bin/env python3
"""GPU experiment: the epoch (count-caching) word kernel against the one-gather-per-step word kernel.

  1. correctness on bcc 32^3: EXACT vs screened instantiation identical, sum(accepted dE) == oracle energy change
  2. throughput on bcc 128^3 (AlTiCrMo, 4 shells, 1000 K) per layout
  3. sampling efficiency: energy relaxation from a random start, E/N after s sweeps, per layout and for the oracle's
     sequential sampler (the reference's proposal distribution) -- attempts/s only count if a sweep does the same work

Writes one JSON line per measurement to stdout (run under gpurun, tee into gpurun_out/)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import brawl_b200 as bw                      # noqa: E402
from oracle import oracle as orc             # noqa: E402  (checker only)

gold = np.load(os.path.join(ROOT, "tests", "golden", "brawl_golden.npz"))
V4 = np.ascontiguousarray(gold["ex_AlTiCrMo_V"][:64])
LAYOUTS = {"word_split(r01)": 3, "epoch2": 5, "epoch4": 0, "epoch8": 4}


def rand_config(n, S, seed):
    rng = np.random.default_rng(seed)
    N = 2 * n ** 3
    spec = np.repeat(np.arange(1, S + 1, dtype=np.int8), -(-N // S))[:N]
    rng.shuffle(spec)
    par = (np.arange(2 * n) & 1).astype(np.int8)
    mask = (par[None, None, :] == par[:, None, None]) & (par[None, :, None] == par[:, None, None])
    g = np.zeros((2 * n, 2 * n, 2 * n), dtype=np.int8)
    g[mask] = spec
    return g


def out(**kw):
    print(json.dumps(kw), flush=True)


def correctness(S=4, T=1000.0):
    n = 32
    V = V4 if S == 4 else np.ascontiguousarray(gold["ex_AlCrFeCoNi_V"][: S * S * 4])
    sysm = orc.System("bcc", n, n, n, S, 4, V)
    g = rand_config(n, S, 5)
    N = sysm.n_atoms
    e0 = sysm.total_energy(g)
    for name, lay in LAYOUTS.items():
        res = []
        for mode in (2, 0):
            dev = bw.Device("bcc", n, n, n, S, 4, V)
            dev.metropolis_set_layout(lay)
            dev.metropolis_set_mode(mode)
            plan = dev.metropolis_plan()
            dev.set_config(g)
            att, acc, dE = dev.metropolis_run(1.0 / (T * bw.K_B_IN_RY), 10 * N, seed=99)
            g1 = dev.get_config().copy()
            res.append((g1, int(att[0]), int(acc[0]), float(dE[0]), plan))
        same = bool(np.array_equal(res[0][0], res[1][0]) and res[0][2] == res[1][2])
        e1 = sysm.total_energy(res[1][0])
        cons = abs((e1 - e0) - res[1][3])
        counts_ok = bool(np.array_equal(np.bincount(res[1][0].ravel(), minlength=6), np.bincount(g.ravel(), minlength=6)))
        out(test="correctness", S=S, T=T, layout=name, plan=res[0][4], att=res[0][1], acc=res[0][2], identical=same,
            conservation_err=cons, dE_screened_minus_exact=res[0][3] - res[1][3], counts_ok=counts_ok)


def throughput():
    n = 128
    g = rand_config(n, 4, 1)
    N = 2 * n ** 3
    beta = 1.0 / (1000.0 * bw.K_B_IN_RY)
    for name, lay in LAYOUTS.items():
        for steps in (0,):
            dev = bw.Device("bcc", n, n, n, 4, 4, V4)
            dev.metropolis_set_layout(lay)
            if steps:
                dev.metropolis_tune((0, 0, 0), steps)
            plan = dev.metropolis_plan()
            dev.set_config(g)
            dev.metropolis_run(beta, 16 * N)
            best = 0.0
            for rep in range(3):
                t0 = time.perf_counter()
                att, acc, dE = dev.metropolis_run(beta, 64 * N)
                dt = time.perf_counter() - t0
                best = max(best, att[0] / dt)
            out(test="throughput", layout=name, plan=plan, swaps_per_s=best, acceptance=float(acc[0]) / float(att[0]),
                launches=dev.metropolis_last_launches())


def relaxation():
    n = 32
    sysm = orc.System("bcc", n, n, n, 4, 4, V4)
    N = sysm.n_atoms
    T = 1000.0
    beta = 1.0 / (T * bw.K_B_IN_RY)
    marks = [1, 2, 4, 8, 16, 32, 64, 128]
    R = 8
    for name, lay in LAYOUTS.items():
        dev = bw.Device("bcc", n, n, n, 4, 4, V4, n_replicas=R)
        dev.metropolis_set_layout(lay)
        dev.set_config(np.stack([rand_config(n, 4, 100 + r) for r in range(R)]))
        curve, done = [], 0
        tot_att = 0
        for m in marks:
            att, acc, dE = dev.metropolis_run(beta, (m - done) * N, seed=7)
            tot_att += int(att[0])
            done = m
            e = dev.total_energy(0, R, exact_order=False) / N
            curve.append((tot_att / N, float(e.mean()), float(e.std(ddof=1) / np.sqrt(R))))
        out(test="relaxation", layout=name, T=T, curve=curve)
    # oracle sequential sampler: 2 chains
    curves = []
    for c in range(2):
        g = rand_config(n, 4, 100 + c)
        mt = orc.MT(seed=900 + c)
        done, cv = 0, []
        for m in marks[:6]:
            sysm.metropolis_trials(g, mt, beta, (m - done) * N)
            done = m
            cv.append((float(m), sysm.total_energy(g) / N))
        curves.append(cv)
    out(test="relaxation", layout="oracle_sequential", T=T, curve=[(a[0], 0.5 * (a[1] + b[1]), abs(a[1] - b[1]) / 2) for a, b in zip(*curves)])


def correctness5():
    correctness(5, 800.0)
    correctness(4, 300.0)


if __name__ == "__main__":
    which = sys.argv[1:] or ["correctness", "throughput", "relaxation"]
    for w in which:
        globals()[w]()
