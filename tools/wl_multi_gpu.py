#!/usr/bin/env python3
"""Wang-Landau with the energy windows sharded over the GPUs of one node.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/wl_multi_gpu.py [--windows 8] [--walkers 16] [--tolerance 5e-5]

Workload: the reference's tests/04_parallel_wang-landau / examples/02 shape (bcc n=4, 128 atoms, 4 species,
6 shells, 512 bins).  Rank 0 prints one JSON line: time to the final ln g(E), trials/s, and the NRMSE
against the reference's golden wl_dos.nc (criterion of tests/ci_test.py: < 1 %).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=4)
    ap.add_argument("--walkers", type=int, default=8)
    ap.add_argument("--overlap", type=float, default=0.25)
    ap.add_argument("--tolerance", type=float, default=5e-5)
    ap.add_argument("--seed", type=int, default=2024)
    ap.add_argument("--out", default=None, help="directory for data/wl_dos.nc, wl_dos_bins.nc, wl_hist.nc (the reference's files)")
    ap.add_argument("--comm", default="abi", choices=["abi", "torch"],
                    help="collectives through the C ABI's NCCL communicator (brawl_cuda_comm_*) or torch.distributed")
    ap.add_argument("--span", action="store_true",
                    help="every GPU holds --walkers walkers of EVERY window; the window average becomes an ncclAllReduce "
                         "(brawl_cuda_wl_set_span): for fewer windows than GPUs")
    ap.add_argument("--performance", type=int, default=4,
                    help="the reference's switch: 0/1 resize windows every f-stage, 2/3 after pre-sampling only, 4 static")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import brawl_b200
    from brawl_b200 import wang_landau as wl
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gold = np.load(os.path.join(ROOT, "tests", "golden", "brawl_golden.npz"))
    p = wl.WLParams(mc_sweeps=100, bins=512, num_windows=args.windows, bin_overlap=args.overlap, tolerance=args.tolerance,
                    flatness=0.90, wl_f=0.05, energy_min=-96, energy_max=0.0, performance=args.performance)
    uid = None
    if world > 1 and args.comm == "abi":                        # rank 0's NCCL id reaches the others through torch's broadcast
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t.copy_(torch.from_numpy(brawl_b200.Device("bcc", 4, 4, 4, 4, 6, gold["t04_V"], device=local).comm_unique_id()))
        dist.broadcast(t, 0)
        uid = t.cpu().numpy()
    drv = wl.WangLandau("bcc", 4, 4, 4, 4, 6, gold["t04_V"], [32] * 4, p, walkers=args.walkers, device=local, rank=rank,
                        world=world, seed=args.seed, torch_device=torch.device("cuda", local),
                        comm=args.comm if world > 1 else "torch", unique_id=uid, span=args.span)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    lng = drv.run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.time() - t0
    trials = drv.comm.all_sum(drv.total_trials)
    if args.out:
        drv.save_wl_data(args.out, lng)
    if rank == 0:
        ref = np.asarray(gold["t04_wl_dos"], dtype=np.float64)
        err = float(np.sqrt(np.mean((ref - lng) ** 2)) / np.mean(np.abs(ref)))
        print(json.dumps({"workload": "WL bcc n=4 4 species 6 shells 512 bins", "n_gpus": world, "windows": args.windows,
                          "walkers_per_window": drv.walkers_total, "span": bool(drv.span), "performance": args.performance, "comm": args.comm if world > 1 else "none",
                          "final_window_widths": (drv.window_indices[:, 1] - drv.window_indices[:, 0] + 1).tolist(),
                          "seconds_to_final_lng": dt, "wl_trials": trials,
                          "wl_trials_per_sec": trials / dt, "sweeps_calls_per_stage": drv.stage_sweeps,
                          "nrmse_vs_reference_golden": err, "pass_reference_criterion": err < 0.01}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
