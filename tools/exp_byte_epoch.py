#!/usr/bin/env python3
"""GPU experiment: the byte-lattice epoch kernels (epoch_byte_metropolis.cuh: fcc, 6-shell bcc) against the
one-gather-per-step byte kernels.

  1. correctness on 32^3 cells: EXACT vs screened instantiation identical, sum(accepted dE) == oracle energy change
  2. throughput on 128^3 cells at 1000 K per layout (0: epochs of 4 steps, 4: epochs of 8, 1: gather per step)
  3. sampling efficiency: energy relaxation after a quench from a random start, E/N after s sweeps per layout

One JSON line per measurement on stdout (run under gpurun, tee into gpurun_out/)."""
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
CASES = [("bcc", 4, 6, "t02_V", 4), ("fcc", 5, 4, "ex_AlCrFeCoNi_V", 5), ("fcc", 2, 4, "ex_FeNi_V", 2), ("fcc", 5, 6, "t01_V", 5)]
LAYOUTS = {"gather_per_step": 1, "epoch4": 0, "epoch8": 4}


def table(key, S0, S, shells):
    return np.ascontiguousarray(gold[key][: S0 * S0 * shells].reshape(shells, S0, S0)[:, :S, :S]).ravel()


def rand_config(lattice, n, S, seed):
    rng = np.random.default_rng(seed)
    par = (np.arange(2 * n) & 1).astype(np.int8)
    if lattice == "bcc":
        mask = (par[None, None, :] == par[:, None, None]) & (par[None, :, None] == par[:, None, None])
    else:
        mask = ((par[None, None, :] + par[None, :, None] + par[:, None, None]) & 1) == 0
    N = int(mask.sum())
    spec = np.repeat(np.arange(1, S + 1, dtype=np.int8), -(-N // S))[:N]
    rng.shuffle(spec)
    g = np.zeros((2 * n, 2 * n, 2 * n), dtype=np.int8)
    g[mask] = spec
    return g


def out(**kw):
    print(json.dumps(kw), flush=True)


def correctness(lattice, S, shells, key, S0, T=800.0):
    n = 32
    V = table(key, S0, S, shells)
    sysm = orc.System(lattice, n, n, n, S, shells, V)
    g = rand_config(lattice, n, S, 5)
    N = sysm.n_atoms
    e0 = sysm.total_energy(g)
    for name, lay in LAYOUTS.items():
        if lay == 1:
            continue
        res = []
        for mode in (2, 0):
            dev = bw.Device(lattice, n, n, n, S, shells, V)
            dev.metropolis_set_layout(lay)
            dev.metropolis_set_mode(mode)
            plan = dev.metropolis_plan()
            dev.set_config(g)
            att, acc, dE = dev.metropolis_run(1.0 / (T * bw.K_B_IN_RY), 10 * N, seed=99)
            g1 = dev.get_config().copy()
            res.append((g1, int(att[0]), int(acc[0]), float(dE[0]), plan))
            dev.close()
        e1 = sysm.total_energy(res[1][0])
        out(test="correctness", lattice=lattice, S=S, shells=shells, layout=name, plan=res[0][4],
            identical=bool(np.array_equal(res[0][0], res[1][0])), att=res[0][1], acc=(res[0][2], res[1][2]),
            dE=(res[0][3], res[1][3]), oracle_dE=e1 - e0, conserved=bool(abs((e1 - e0) - res[1][3]) < 1e-10 * abs(e1 - e0) + 1e-11),
            counts_ok=bool(np.array_equal(np.bincount(res[0][0].ravel(), minlength=S + 1), np.bincount(g.ravel(), minlength=S + 1))))


def throughput(lattice, S, shells, key, S0, T=1000.0, n=128):
    V = table(key, S0, S, shells)
    g = rand_config(lattice, n, S, 1)
    for name, lay in LAYOUTS.items():
        dev = bw.Device(lattice, n, n, n, S, shells, V)
        dev.metropolis_set_layout(lay)
        N = dev.n_atoms
        dev.set_config(g)
        beta = 1.0 / (T * bw.K_B_IN_RY)
        plan = dev.metropolis_plan()
        dev.metropolis_run(beta, 8 * N)
        best = 0.0
        for _ in range(3):
            t0 = time.perf_counter()
            att, acc, dE = dev.metropolis_run(beta, 32 * N)
            best = max(best, att[0] / (time.perf_counter() - t0))
        out(test="throughput", lattice=lattice, S=S, shells=shells, layout=name, swaps_per_s=best,
            acceptance=float(acc[0]) / float(att[0]), plan=plan)
        dev.close()


def relaxation(lattice, S, shells, key, S0, T=1000.0, n=32):
    """E/N after s sweeps from a random start (mean over 4 replicas), per layout, and for the oracle's sequential sampler"""
    V = table(key, S0, S, shells)
    R = 4
    gs = np.stack([rand_config(lattice, n, S, 10 + r) for r in range(R)])
    beta = 1.0 / (T * bw.K_B_IN_RY)
    points = [1, 2, 4, 8, 16, 32]
    for name, lay in LAYOUTS.items():
        dev = bw.Device(lattice, n, n, n, S, shells, V, n_replicas=R)
        dev.metropolis_set_layout(lay)
        N = dev.n_atoms
        dev.set_config(gs)
        done, es, sw = 0, [], []
        for s in points:
            att, acc, dE = dev.metropolis_run(beta, (s - done) * N, seed=5 + s)
            done = s
            sw.append(float(np.sum(att)) / (R * N) + (sw[-1] if sw else 0.0))
            es.append(float(np.mean(dev.total_energy(0, R, exact_order=False))) / N)
        out(test="relaxation", lattice=lattice, S=S, shells=shells, layout=name, sweeps=sw, E=es)
        dev.close()
    sysm = orc.System(lattice, n, n, n, S, shells, V)
    N = sysm.n_atoms
    g = gs[0].copy()
    mt = orc.MT(seed=11)
    done, es = 0, []
    for s in points[:5]:
        sysm.metropolis_trials(g, mt, beta, (s - done) * N)
        done = s
        es.append(sysm.total_energy(g) / N)
    out(test="relaxation", lattice=lattice, S=S, shells=shells, layout="oracle_sequential", sweeps=points[:5], E=es)


if __name__ == "__main__":
    what = sys.argv[1:] or ["correctness", "throughput", "relaxation"]
    for c in CASES:
        if "correctness" in what:
            correctness(*c)
        if "throughput" in what:
            throughput(*c)
    if "relaxation" in what:
        relaxation(*CASES[0])
        relaxation(*CASES[1])
