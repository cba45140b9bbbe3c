import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import brawl_b200 as bw
from oracle import oracle as orc
gold = np.load("tests/golden/brawl_golden.npz")
V = gold["ex_AlTiCrMo_V"][:64]
n = 32
sysm = orc.System("bcc", n, n, n, 4, 4, V)
mt = orc.MT(seed=3)
c, q = sysm.quotas(conc=[0.25] * 4)
g = sysm.initial_setup(mt, c, q)
N = sysm.n_atoms
for T in (1200.0, 2500.0):
    for layout, mode in ((False, 2), (False, 0), (True, 1)):
        dev = bw.Device("bcc", n, n, n, 4, 4, V)
        dev.metropolis_set_layout(layout); dev.metropolis_set_mode(mode)
        dev.set_config(g)
        beta = 1.0 / (T * bw.K_B_IN_RY)
        tr = []
        for k in range(30):
            att, acc, dE = dev.metropolis_run(beta, 20 * N, seed=100 + k)
            tr.append(dev.total_energy()[0] / N)
        print(T, layout, mode, dev.metropolis_plan()["use_box"], "acc %.4f" % (acc[0] / att[0]), " ".join("%.6f" % e for e in tr[::3]), flush=True)
