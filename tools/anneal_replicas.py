#!/usr/bin/env python3
"""BASELINE config 5 through the annealing driver: REPLICAS (default 1024) quinary replicas of bcc 32^3 per GPU cooled
over a T ladder, energies sampled from the batched total_energy kernel and SRO from the batched radial-counts kernel.

    python tools/anneal_replicas.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/anneal_replicas.py                                      # N GPUs, REPLICAS per GPU (weak scaling)

Rank 0 prints the whole-run attempted swaps/s over all GPUs (sampling included; max wall time over ranks between two
barriers) and the SRO-vs-T table averaged over all chains of all ranks (the reference's av_* outputs, comms.F90:122-160).
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brawl_b200 import replica_annealing as ra


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    tdev = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        tdev = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=tdev)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "brawl_golden.npz"))
    V = gold["ex_AlCrFeCoNi_V"][:100]
    R, n = int(os.environ.get("REPLICAS", 1024)), 32
    N = 2 * n ** 3
    counts = [N // 5 + (1 if s < N % 5 else 0) for s in range(5)]
    drv = ra.ReplicaAnnealing("bcc", n, n, n, 5, 4, V, counts, n_replicas=R, T=3000.0, T_steps=6, delta_T=-500.0,
                              n_mc_steps=40 * N, n_sample_steps=10 * N, n_burn_in_steps=20 * N, burn_in_start=True,
                              burn_in=True, n_sample_steps_asro=20 * N, wc_range=3, device=local, rank=rank, world=world,
                              seed=0x42726157 + rank, torch_device=tdev)
    drv.run()                                   # warm-up (plans, allocations)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    per, av = drv.run()
    dt = time.perf_counter() - t0
    dt = float(drv.comm.all_gather(np.array([dt])).max())
    attempted = drv.comm.all_sum(drv.attempted)
    var = drv.comm.all_gather((per["energies_of_T"] ** 2).sum(axis=0)).sum(axis=0) / av["n_chains"] - av["energies_of_T"] ** 2
    sem = np.sqrt(np.maximum(var, 0.0) / av["n_chains"])
    if rank == 0:
        print("%d GPU(s) x %d replicas x bcc %d^3 quinary, 6 temperatures x (20 + 40) sweeps, 4 energy + 2 SRO samples per T: "
              "%.2f s, %.3g attempted swaps/s incl. sampling" % (world, R, n, dt, attempted / dt))
        a = ra.warren_cowley(av["rho_of_T"][:, 1:], [c / N for c in counts], [8, 6])
        for j, T in enumerate(av["temperature"]):
            print("T %6.0f K  <E> %9.5f mRy/atom  acc %.3f  alpha1(Al-Al) %+.4f  alpha1(Al-Ni) %+.4f  sem(E) %.2e" % (
                T, 1e3 * av["energies_of_T"][j], av["acceptance_of_T"][j], a[j, 0, 0, 0], a[j, 0, 4, 0], 1e3 * sem[j]))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
