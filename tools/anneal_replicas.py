"""BASELINE config 5 through the annealing driver: 1024 quinary replicas of bcc 32^3 per GPU cooled over a T ladder,
energies sampled from the batched total_energy kernel and SRO from the batched radial-counts kernel.  Prints the
whole-run attempted swaps/s (sampling included) and the SRO-vs-T table."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brawl_b200 import replica_annealing as ra

gold = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "brawl_golden.npz"))
V = gold["ex_AlCrFeCoNi_V"][:100]
R, n = int(os.environ.get("REPLICAS", 1024)), 32
N = 2 * n ** 3
counts = [N // 5 + (1 if s < N % 5 else 0) for s in range(5)]
drv = ra.ReplicaAnnealing("bcc", n, n, n, 5, 4, V, counts, n_replicas=R, T=3000.0, T_steps=6, delta_T=-500.0, n_mc_steps=40 * N,
                          n_sample_steps=10 * N, n_burn_in_steps=20 * N, burn_in_start=True, burn_in=True,
                          n_sample_steps_asro=20 * N, wc_range=3)
drv.run()                                   # warm-up (plans, allocations)
t0 = time.perf_counter()
per, av = drv.run()
dt = time.perf_counter() - t0
print("replicas %d x bcc %d^3 quinary, 6 temperatures x (20 + 40) sweeps, 4 energy + 2 SRO samples per T: %.2f s, %.3g attempted swaps/s incl. sampling"
      % (R, n, dt, drv.attempted / dt))
a = ra.warren_cowley(av["rho_of_T"][:, 1:], [c / N for c in counts], [8, 6])
for j, T in enumerate(av["temperature"]):
    print("T %6.0f K  <E> %9.5f mRy/atom  acc %.3f  alpha1(Al-Al) %+.4f  alpha1(Al-Ni) %+.4f  sem(E) %.2e" % (
        T, 1e3 * av["energies_of_T"][j], av["acceptance_of_T"][j], a[j, 0, 0, 0], a[j, 0, 4, 0],
        1e3 * per["energies_of_T"][:, j].std(ddof=1) / np.sqrt(R)))
