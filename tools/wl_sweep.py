#!/usr/bin/env python3
"""Wang-Landau time to the final ln g and NRMSE against the reference's golden wl_dos.nc over window counts, walkers per
window and seeds, on one GPU (bench.wl_time_to_flatness).  usage: wl_sweep.py W1,W2,.. K1,K2,.. seed1,seed2,.."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

windows = [int(v) for v in sys.argv[1].split(",")]
walkers = [int(v) for v in sys.argv[2].split(",")]
seeds = [int(v) for v in sys.argv[3].split(",")]
for W in windows:
    for K in walkers:
        for s in seeds:
            b = bench.wl_time_to_flatness(0, 1, 0, W, K, seed=s, performance=int(os.environ.get("WL_PERF", "4")))
            print(json.dumps({"windows": W, "walkers": K, "seed": s, "seconds": round(b["value"], 3),
                              "nrmse": round(b["nrmse_vs_reference_golden"], 5), "calls": sum(b["sweeps_calls_per_stage"]),
                              "host": b["host_seconds_per_rank"][0]}), flush=True)
