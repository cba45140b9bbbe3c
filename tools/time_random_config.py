"""Time brawl_cuda_random_config against the host route (numpy shuffle + set_config) it replaces."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import brawl_b200
from brawl_b200 import wang_landau as wl

gold = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "brawl_golden.npz"))
for name, n, S, R, V in (("1024 x 32^3 bcc quinary", 32, 5, 1024, gold["ex_AlCrFeCoNi_V"][:100]),
                         ("1 x 128^3 bcc quaternary", 128, 4, 1, gold["ex_AlTiCrMo_V"][:64])):
    dev = brawl_b200.Device("bcc", n, n, n, S, 4, V, n_replicas=R)
    na = dev.n_atoms
    counts = [na // S + (1 if s < na % S else 0) for s in range(S)]
    dev.random_config(counts, 0, R, seed=1)                        # warm-up (allocations)
    t0 = time.perf_counter()
    for k in range(3):
        dev.random_config(counts, 0, R, seed=2 + k)
    t_dev = (time.perf_counter() - t0) / 3
    rng = np.random.default_rng(0)
    m = min(R, 16)
    t0 = time.perf_counter()
    for r in range(m):
        dev.set_config(wl.random_configuration("bcc", n, n, n, counts, rng), r, 1)
    t_host = (time.perf_counter() - t0) / m * R
    print("%s: device %.2f ms (%.2f Gsites/s), host shuffle + set_config %.1f ms" % (name, t_dev * 1e3, na * R / t_dev / 1e9, t_host * 1e3))
