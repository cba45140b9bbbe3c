#!/usr/bin/env python3
"""Latency of the C ABI's NCCL collectives (brawl_cuda_comm_*) under torchrun: prints per-call milliseconds on rank 0."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import brawl_b200

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gold = np.load(os.path.join(ROOT, "tests", "golden", "brawl_golden.npz"))
dev = brawl_b200.Device("bcc", 4, 4, 4, 4, 6, gold["t04_V"], device=local, n_replicas=128)
t = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    t.copy_(torch.from_numpy(dev.comm_unique_id()))
dist.broadcast(t, 0)
dev.comm_create(world, rank, t.cpu().numpy())
dev.wl_init(512, np.linspace(-1.0, 0.0, 513), 16)
a = np.arange(129.0)
for name, fn in (("comm_allgather(129)", lambda: dev.comm_allgather(a)), ("comm_allreduce(16)", lambda: dev.comm_allreduce(a[:16])),
                 ("wl_allgather_lng(8x512)", lambda: dev.wl_allgather_lng(world)),
                 ("exchange_replica(rank^1)", lambda: dev.exchange_replica(3, rank ^ 1) or dev.synchronize()),
                 ("exchange_replicas(ring)", lambda: dev.exchange_replicas([3, 4], [(rank + 1) % world, (rank - 1) % world]) or dev.synchronize())):
    for _ in range(5):
        fn()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(200):
        fn()
    dt = (time.perf_counter() - t0) / 200
    if rank == 0:
        print("%-28s %.3f ms per call" % (name, 1e3 * dt), flush=True)
x = torch.zeros(129, dtype=torch.float64, device="cuda")
out = [torch.empty_like(x) for _ in range(world)]
for _ in range(5):
    dist.all_gather(out, x)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    dist.all_gather(out, x)
    torch.cuda.synchronize()
if rank == 0:
    print("torch all_gather(129) + sync     %.3f ms per call" % (1e3 * (time.perf_counter() - t0) / 200))
dist.destroy_process_group()
