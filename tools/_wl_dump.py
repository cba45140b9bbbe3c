import os, sys, json, numpy as np
sys.path.insert(0, os.getcwd())
from brawl_b200 import wang_landau as wl
gold = np.load("tests/golden/brawl_golden.npz")
out = {}
for W, seed in ((64, 2024), (64, 1), (64, 7), (8, 2024), (32, 2024)):
    p = wl.WLParams(mc_sweeps=100, bins=512, num_windows=W, bin_overlap=0.25, tolerance=5e-5, flatness=0.90, wl_f=0.05, energy_min=-96, energy_max=0.0, performance=4)
    drv = wl.WangLandau("bcc", 4, 4, 4, 4, 6, gold["t04_V"], [32] * 4, p, walkers=16, device=0, seed=seed)
    lng = drv.run()
    out["lng_%d_%d" % (W, seed)] = lng
    out["wi_%d" % W] = np.array(drv.window_indices)
    ref = np.asarray(gold["t04_wl_dos"], dtype=np.float64)
    print(W, seed, float(np.sqrt(np.mean((ref - lng) ** 2)) / np.mean(np.abs(ref))), flush=True)
np.savez("gpurun_out/wl_dump.npz", **out)
