#!/usr/bin/env python3
"""Tiny production-Metropolis run for `compute-sanitizer --tool racecheck`: if two simultaneous
trials of one step ever touched interacting sites, racecheck would report a shared-memory hazard
between the box store of one thread and the neighbour gather of another.
    compute-sanitizer --tool racecheck --racecheck-report all python tools/racecheck_box.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import brawl_b200

gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "brawl_golden.npz"))
rng = np.random.default_rng(0)
# layout: 0 automatic (epoch kernels: word lattice for bcc-4, byte lattice for fcc / bcc-6), 1 byte-lattice gather-per-step
# kernels, 2 / 3 the round-1 word kernels without / with the two-warp-group split
for lattice, n, S, shells, key, nbr, mode, generic, layout in (
        ("bcc", 32, 4, 4, "ex_AlTiCrMo_V", False, 2, False, 0),      # epoch kernel: mbarrier epochs, TMA box load, dependent launch
        ("bcc", 32, 5, 4, "ex_AlCrFeCoNi_V", False, 2, False, 0),    # epoch kernel, 5 species
        ("bcc", 32, 4, 4, "ex_AlTiCrMo_V", False, 0, False, 0),      # epoch kernel, EXACT instantiation
        ("fcc", 32, 5, 4, "ex_AlCrFeCoNi_V", False, 2, False, 0),    # byte-lattice epoch kernel, fcc 4 shells
        ("bcc", 32, 4, 6, "t02_V", False, 2, False, 0),              # byte-lattice epoch kernel, bcc 6 shells
        ("fcc", 32, 2, 4, "ex_FeNi_V", False, 0, False, 0),          # byte-lattice epoch kernel, EXACT, binary
        ("bcc", 32, 4, 4, "ex_AlTiCrMo_V", False, 2, False, 3),      # round-1 word kernel, two warp groups (named barriers)
        ("bcc", 32, 4, 4, "ex_AlTiCrMo_V", False, 1, False, 1),
        ("fcc", 32, 5, 4, "ex_AlCrFeCoNi_V", False, 1, False, 1),
        ("bcc", 16, 4, 6, "t02_V", False, 0, True, 1),
        ("bcc", 16, 4, 4, "ex_AlTiCrMo_V", True, 0, True, 1)):
    V = gold[key][: S * S * shells]
    par = np.arange(2 * n) & 1
    mask = ((par[None, None, :] == par[:, None, None]) & (par[None, :, None] == par[:, None, None])) if lattice == "bcc" else \
        (((par[None, None, :] + par[None, :, None] + par[:, None, None]) & 1) == 0)
    g = np.zeros((2 * n,) * 3, dtype=np.int8)
    g[mask] = rng.integers(1, S + 1, size=int(mask.sum()))
    dev = brawl_b200.Device(lattice, n, n, n, S, shells, V)
    dev.metropolis_set_mode(mode)
    dev.metropolis_set_layout(layout)
    dev.metropolis_tune((0, 0, 0), -7 if generic else 6)        # 6 steps per phase; negative: generic kernel
    dev.set_config(g)
    plan = dev.metropolis_plan(nbr)
    att, acc, dE = dev.metropolis_run(1.0 / (800.0 * brawl_b200.K_B_IN_RY), 1, nbr_swap=nbr)   # exactly one phase
    e_tile = dev.total_energy(exact_order=False)[0]            # tiled energy kernel (shared-memory tile + halo)
    print(lattice, n, shells, "nbr" if nbr else "lattice", "mode", mode, "layout", layout, "kind", plan["use_box"],
          "groups", plan["warp_groups"], int(att[0]), int(acc[0]), e_tile)
print("done")
