#!/usr/bin/env python3
"""A few `sweeps` calls of the device-resident Wang-Landau loop (8 windows x 16 walkers, the bench shape) for ncu."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brawl_b200 import wang_landau as wl  # noqa: E402

gold = np.load(os.path.join(ROOT, "tests", "golden", "brawl_golden.npz"))
p = wl.WLParams(mc_sweeps=100, bins=512, num_windows=8, bin_overlap=0.25, tolerance=5e-5, flatness=0.90, wl_f=0.05,
                energy_min=-96, energy_max=0.0, performance=4)
drv = wl.WangLandau("bcc", 4, 4, 4, 4, 6, gold["t04_V"], [32] * 4, p, walkers=16, seed=2024)
drv.enter_energy_windows()
for _ in range(8):
    drv._sweeps(0.05)
print("ok", drv.energies[:4])
