#!/usr/bin/env python3
"""bench.py -- attempted swaps per second of the BraWl atom-swap hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # CUDA arm (libbrawl_cuda.so)
    python bench.py --impl reference --gpus N ...            # CPU arm: the reference algorithm on host cores

Workload (BASELINE.json configs[1]): AlTiCrMo bcc Metropolis on a 128^3 lattice (4 194 304 atoms,
reference grid 256^3 = 16 MiB int8), 4 species at 0.25, first 4 shells of AlTiCrMo.vij (Z = 50),
whole-lattice swaps (nbr_swap = F), T = 1000 K, synthetic random start from initial_setup
semantics.  One "step" = --sweeps lattice sweeps (default 128) = sweeps * N_atoms attempted swaps.

Timing: CUDA events on the launching stream around each step, L2 flushed (512 MiB memset) before
every timed step, W warm-up steps, max over ranks.  `value` has the lattice resident in HBM;
`e2e` runs the same step through the host-facing C-ABI call sequence a Fortran driver would make
(pinned host config -> set_config -> metropolis_run -> get_config + total_energy), copies timed.
N > 1: the single chain does not shard ("replicas only"): every rank runs its own
independent 128^3 replica on its own GPU, no data-path collective; value = total attempts / max time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CELLS = 128
T_KELVIN = 1000.0
B_ALG = 104          # algorithmic bytes per attempted swap, bcc 4 shells: 2Z+4
METRIC = "attempted_swaps_per_sec"


# workload -> (lattice, species of the stored table, species, shells, golden key, name)
OTHER_LATTICES = {"bcc6": ("bcc", 4, 4, 6, "t02_V", "AlTiCrMo"), "fcc4": ("fcc", 5, 5, 4, "ex_AlCrFeCoNi_V", "AlCrFeCoNi"),
                  "feni": ("fcc", 2, 2, 4, "ex_FeNi_V", "FeNi"), "fcc6": ("fcc", 5, 5, 6, "t01_V", "AlCrFeCoNi")}


def load_V():
    """First 4 shell blocks of examples/02_wang-landau_AlTiCrMo/AlTiCrMo.vij (committed fixture)."""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "brawl_golden.npz"))
    return np.ascontiguousarray(gold["ex_AlTiCrMo_V"][: 4 * 4 * 4])


def synthetic_config(n, S, rank, lattice="bcc"):
    """Random equiatomic start written directly on the lattice sites (numpy; same distribution as
    initial_setup: a uniformly random arrangement of the species multiset)."""
    rng = np.random.default_rng(110179 + 11 * rank)
    par = (np.arange(2 * n) & 1).astype(np.int8)
    if lattice == "bcc":
        mask = (par[None, None, :] == par[:, None, None]) & (par[None, :, None] == par[:, None, None])
    else:
        mask = ((par[None, None, :] + par[None, :, None] + par[:, None, None]) & 1) == 0
    N = int(mask.sum())
    spec = np.repeat(np.arange(1, S + 1, dtype=np.int8), -(-N // S))[:N]
    rng.shuffle(spec)
    g = np.zeros((2 * n, 2 * n, 2 * n), dtype=np.int8)
    g[mask] = spec
    return g


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (rows are time-stamped on
    arrival and filtered to the [mark_start, mark_end] window)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            t_end = time.time() + 3.0
            while not self.rows and time.time() < t_end:      # wait until nvidia-smi is actually sampling
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = self.t0 or 0.0
        t1 = (self.t1 or time.time()) + 0.06
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for t, r in self.rows[-2:] if len(r) >= 7]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None}


def cpu_reference_rate(n_threads, trials_per_thread, n=N_CELLS, seed0=0, S=4, V=None, temp=T_KELVIN):
    """The reference algorithm (oracle C restatement, validated bit-exactly against the reference's
    goldens) on host cores: one independent replica per thread, the reference's own Metropolis
    parallel model (src/comms.F90:122-160).  Returns (attempted swaps/s over all threads, seconds)."""
    from oracle import oracle          # CPU baseline legs only: the oracle is the thing being timed here
    V = load_V() if V is None else V
    osys = oracle.System("bcc", n, n, n, S, 4, V)
    beta = 1.0 / (temp * oracle.K_B_IN_RY)
    grids = [synthetic_config(n, S, seed0 + t) for t in range(n_threads)]
    mts = [oracle.MT(rank=seed0 + t) for t in range(n_threads)]
    for t in range(n_threads):                         # warm the caches / page in
        osys.metropolis_trials(grids[t], mts[t], beta, 20000)
    ths = [threading.Thread(target=osys.metropolis_trials, args=(grids[t], mts[t], beta, trials_per_thread))
           for t in range(n_threads)]                  # ctypes releases the GIL
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    return n_threads * trials_per_thread / dt, dt


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_name(n):
    """config.workload of the headline configuration -- the same string in both arms (--impl cuda / reference)."""
    return ("AlTiCrMo bcc %d^3 (%d atoms), 4 species @0.25, 4 shells (Z=50), Metropolis whole-lattice swaps, T=%g K; "
            "replicas only for N>1 (one chain per GPU)" % (n, 2 * n ** 3, T_KELVIN))


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_thread = args.ref_trials
    vals, times = [], []
    for s in range(args.warmup + args.steps):
        v, dt = cpu_reference_rate(cores, per_thread, seed0=s * cores)
        if s >= args.warmup:
            vals.append(v); times.append(dt)
    value = float(np.mean(vals))
    sample = "%d threads x %d trials per step on private 128^3 bcc replicas (T=%g K)" % (cores, per_thread, T_KELVIN)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "swaps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(N_CELLS), "n_atoms": 2 * N_CELLS ** 3, "step": sample},
        "cpu_baseline": {"value": value, "unit": "swaps/s", "cores": cores, "kind": "port",
                         "sample": sample + "; C restatement of the reference (no Fortran toolchain in the image), "
                                            "bit-exact vs reference goldens"},
        "e2e": {"value": value, "unit": "swaps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def wl_time_to_flatness(rank, world, local_rank, windows_per_gpu=8, walkers=16, tolerance=5e-5, seed=2024, performance=4):
    """BASELINE.json metric, second half: Wang-Landau time to the final ln g(E) on the reference's regression / example
    shape (tests/04_parallel_wang-landau, examples/02: bcc n=4, 128 atoms, AlTiCrMo 6 shells, 512 bins in [-96, 0]
    meV/atom, overlap 0.25, flatness 0.9, f 0.05 -> tolerance), energy windows sharded over the GPUs (8 per GPU), every
    collective through the C ABI's NCCL communicator.  Checked against the reference's golden wl_dos.nc with its own
    criterion (NRMSE < 1 %, tests/ci_test.py:42-50).  Returns a dict (rank 0) -- the `extra.wl` block."""
    import torch
    import torch.distributed as dist
    from brawl_b200 import wang_landau as wl
    gold = np.load(os.path.join(ROOT, "tests", "golden", "brawl_golden.npz"))
    # the reference's load balancer refuses windows narrower than w_min * bins = 10 bins (mpi_window_optimise,
    # src/wang-landau.F90:1275): at most 51 windows on 512 bins, i.e. 6 per GPU at 8 GPUs (8 per GPU up to 4 GPUs)
    windows_per_gpu = max(1, min(windows_per_gpu, (512 // 10) // world))
    p = wl.WLParams(mc_sweeps=100, bins=512, num_windows=windows_per_gpu * world, bin_overlap=0.25, tolerance=tolerance,
                    flatness=0.90, wl_f=0.05, energy_min=-96, energy_max=0.0, performance=performance)
    uid = None
    if world > 1:
        import brawl_b200
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            probe = brawl_b200.Device("bcc", 4, 4, 4, 4, 6, gold["t04_V"], device=local_rank)
            t.copy_(torch.from_numpy(probe.comm_unique_id()))
        dist.broadcast(t, 0)
        uid = t.cpu().numpy()
    drv = wl.WangLandau("bcc", 4, 4, 4, 4, 6, gold["t04_V"], [32] * 4, p, walkers=walkers, device=local_rank, rank=rank,
                        world=world, seed=seed, comm="abi" if world > 1 else "torch", unique_id=uid)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    lng = drv.run()
    drv.dev.synchronize()
    dt = time.perf_counter() - t0
    trials = float(drv.total_trials)
    timing = [dict(drv.timing)]
    if world > 1:
        tv = torch.tensor([drv.timing[k] for k in sorted(drv.timing)], dtype=torch.float64, device="cuda")
        tl = [torch.empty_like(tv) for _ in range(world)]
        dist.all_gather(tl, tv)
        timing = [dict(zip(sorted(drv.timing), (round(float(x), 4) for x in t))) for t in tl]
        v = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        c = torch.tensor([trials], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        dt, trials = float(v[0]), float(c[0])
    ref = np.asarray(gold["t04_wl_dos"], dtype=np.float64)
    err = float(np.sqrt(np.mean((ref - lng) ** 2)) / np.mean(np.abs(ref)))
    # ln g is normalised to its minimum (the ground-state bin, the hardest to sample): most of the NRMSE is the resulting
    # constant offset; the shape error is what remains when the mean difference is removed
    err_shape = float(np.sqrt(np.mean((ref - lng - np.mean(ref - lng)) ** 2)) / np.mean(np.abs(ref)))
    return {"metric": "wl_seconds_to_final_lng", "value": dt, "unit": "s", "higher_is_better": False, "n_gpus": world,
            "workload": "Wang-Landau bcc n=4 (128 atoms) AlTiCrMo 6 shells, 512 bins, f 0.05 -> %g, flatness 0.9, overlap 0.25" % tolerance,
            "windows": p.num_windows, "walkers_per_window": walkers, "wl_trials": trials, "wl_trials_per_sec": trials / dt,
            "sweeps_calls_per_stage": drv.stage_sweeps, "nrmse_vs_reference_golden": err, "pass_reference_criterion": err < 0.01,
            "nrmse_offset_removed": err_shape,
            "collectives": "C ABI NCCL (brawl_cuda_comm_*)" if world > 1 else "none (one GPU)",
            "host_seconds_per_rank": timing}


def ns_walk_rate(rank, world, local_rank, n_runs=2048, n_iter=100):
    """BASELINE configs[3]: nested sampling on the reference's example shape (examples/03: fcc 3x3x3 = 108 atoms, AlCrFeCoNi,
    4 shells, K = 100 walkers, 500 walk steps per iteration), `n_runs` independent runs per GPU advancing in lock-step: per
    iteration one batched clone launch + one walk kernel (one warp per walking clone).  Returns the `extra.ns` block."""
    import torch
    import torch.distributed as dist
    from brawl_b200 import nested_sampling as ns
    gold = np.load(os.path.join(ROOT, "tests", "golden", "brawl_golden.npz"))
    V = np.ascontiguousarray(gold["t03_V"][: 5 * 5 * 4])
    p = ns.NSParams(n_walkers=100, n_steps=500, n_iter=n_iter)
    drv = ns.NestedSampling("fcc", 3, 3, 3, 5, 4, V, [21, 21, 21, 21, 24], p, n_runs=n_runs, device=local_rank, seed=7 + rank)
    drv.initialise()
    drv.run(n_iter=10)                                         # warm-up
    drv.dev.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    culled = drv.run(n_iter=n_iter)
    drv.dev.synchronize()
    dt = time.perf_counter() - t0
    trials = float(n_runs) * n_iter * p.n_steps                # lower bound: runs with few acceptances walk longer (:129-144)
    if world > 1:
        v = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        dt, trials = float(v[0]), trials * world
    return {"metric": "ns_walk_trials_per_sec", "value": trials / dt, "unit": "trials/s", "higher_is_better": True, "n_gpus": world,
            "workload": "nested sampling, fcc 3x3x3 (108 atoms) AlCrFeCoNi 4 shells, K=100 walkers, 500 walk steps per iteration; "
                        "%d independent runs per GPU in lock-step, %d iterations timed" % (n_runs, n_iter),
            "seconds": dt, "iterations_per_sec_per_run": n_iter / dt, "runs_per_gpu": n_runs,
            "ceilings_monotone": bool(np.all(np.diff(culled, axis=1) <= 0))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--sweeps", type=int, default=128, help="lattice sweeps (N_atoms attempted swaps each) per step")
    ap.add_argument("--ref-trials", type=int, default=1500000, help="--impl reference: trials per thread per step")
    ap.add_argument("--cpu-trials", type=int, default=3000000, help="cpu_baseline leg: trials per thread")
    ap.add_argument("--box", default="", help="override box extents, e.g. 64,64,32")
    ap.add_argument("--steps-per-phase", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layout", type=int, default=0, choices=[0, 1, 2, 3, 4, 5],
                    help="0 automatic (epoch kernel), 1 byte-lattice kernels only, 2 / 3 the one-gather-per-step word kernels "
                         "without / with the warp-group split (3 = the round-1 default), 4 / 5 other epoch lengths (A/B runs)")
    ap.add_argument("--n-cells", type=int, default=N_CELLS)
    ap.add_argument("--workload", default="chain", choices=["chain", "replicas", "wl"],
                    help="chain: BASELINE configs[1] (headline); replicas: configs[4], R x 32^3 bcc AlCrFeCoNi per GPU; "
                         "wl: configs[2], Wang-Landau time to the final ln g only")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra.* blocks (WL time-to-flatness etc.)")
    ap.add_argument("--wl-walkers", type=int, default=16)
    ap.add_argument("--wl-windows-per-gpu", type=int, default=8)
    ap.add_argument("--replicas", type=int, default=1024)
    ap.add_argument("--dE-mode", type=int, default=2, choices=[0, 1, 2],
                    help="2 (library default): word-lattice kernel, integer-count screening with fixed-point dp4a dE, "
                         "reference association recomputed inside the guard band (decision-identical); 1: the same "
                         "screening on the byte lattice; 0: reference association for every trial")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = max(args.warmup, 3) if os.environ.get("BENCH_ALLOW_SHORT_WARMUP") is None else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import brawl_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CUDA arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # NCCL prints its version line to stdout: keep stdout = one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.workload == "wl":
        blk = wl_time_to_flatness(rank, world, local_rank, args.wl_windows_per_gpu, args.wl_walkers)
        if rank == 0:
            print(json.dumps(blk))
        if world > 1:
            dist.destroy_process_group()
        return

    n = args.n_cells
    gold = np.load(os.path.join(ROOT, "tests", "golden", "brawl_golden.npz"))
    stream = torch.cuda.Stream()            # non-default stream shared by torch events and the library
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_device(workload, layout=0, nbr_swap=False):
        """-> (dev, beta, n, R, S, V, description)"""
        if workload == "replicas":
            # BASELINE configs[4]: R independent AlCrFeCoNi replicas (32^3 bcc, 65 536 atoms each) per GPU, first 4 shell
            # blocks of fcc_al_1.00_crfeconi.vij as a synthetic bcc table, annealing ladder 3000 -> 100 K over the replicas
            nn, R, S = 32, args.replicas, 5
            V = np.ascontiguousarray(gold["ex_AlCrFeCoNi_V"][: 5 * 5 * 4])
            base = synthetic_config(nn, 5, rank)
            g0 = np.ascontiguousarray(np.broadcast_to(base, (R,) + base.shape))
            beta = 1.0 / (np.linspace(3000.0, 100.0, R) * brawl_b200.K_B_IN_RY)
            desc = ("%d independent AlCrFeCoNi replicas per GPU, bcc 32^3 (65536 atoms each), 5 species @0.2, 4 shells (Z=50), "
                    "Metropolis whole-lattice swaps, T ladder 3000->100 K" % R)
        elif workload in OTHER_LATTICES:
            # the geometries without a dense-set word kernel: byte-lattice epoch kernels (epoch_byte_metropolis.cuh)
            lattice, S0, S, shells, key, what = OTHER_LATTICES[workload]
            nn, R = n, 1
            V = np.ascontiguousarray(gold[key][: S0 * S0 * shells].reshape(shells, S0, S0)[:, :S, :S]).ravel()
            g0 = synthetic_config(nn, S, rank, lattice)
            beta = None
            Z = {("bcc", 6): 64, ("fcc", 4): 54, ("fcc", 6): 86}[(lattice, shells)]
            desc = "%s %s %d^3 (%d atoms), %d species equiatomic, %d shells (Z=%d), Metropolis whole-lattice swaps" % (
                what, lattice, nn, (2 if lattice == "bcc" else 4) * nn ** 3, S, shells, Z)
            dev = brawl_b200.Device(lattice, nn, nn, nn, S, shells, V, device=local_rank, n_replicas=R)
            dev.b_alg = 2 * Z + 4
        else:
            nn, R, S = n, 1, 4
            V = load_V()
            g0 = synthetic_config(nn, 4, rank)
            beta = None
            desc = None
        if workload not in OTHER_LATTICES:
            dev = brawl_b200.Device("bcc", nn, nn, nn, S, 4, V, device=local_rank, n_replicas=R)
        dev.metropolis_set_mode(args.dE_mode)
        if layout:
            dev.metropolis_set_layout(layout)
        dev.set_stream(stream.cuda_stream)
        if args.box:
            dev.metropolis_tune(tuple(int(v) for v in args.box.split(",")), args.steps_per_phase)
        elif args.steps_per_phase:
            dev.metropolis_tune((0, 0, 0), args.steps_per_phase)
        dev.set_config(g0)
        return dev, beta, nn, R, S, V, desc

    def timed(dev, beta, trials_per_step, steps, warmup, nbr_swap=False, sampler=None):
        """W untimed + K timed steps, CUDA events on the launching stream around each step, L2 flushed before each;
        -> (attempts per replica, total ms, launches, acceptance)"""
        for _ in range(warmup):
            dev.metropolis_enqueue(beta, trials_per_step, nbr_swap=nbr_swap)
        dev.metropolis_counters(reset=True)
        barrier()
        if sampler:
            sampler.start()
            barrier()
            sampler.mark_start()
        times, attempts, launches = [], 0, 0
        for _ in range(steps):
            flush.zero_()                                   # L2 flush, outside the timed events
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            planned, nl = dev.metropolis_enqueue(beta, trials_per_step, nbr_swap=nbr_swap)
            ev1.record(stream)
            ev1.synchronize()
            times.append(ev0.elapsed_time(ev1))
            attempts += planned
            launches += nl
        barrier()
        if sampler:
            sampler.mark_end()
        att, acc, dE = dev.metropolis_counters(reset=True)
        assert att[0] == attempts, (att, attempts)      # planned == counted on the device
        return attempts, float(np.sum(times)), launches, float(acc.sum()) / max(1.0, float(att.sum()))

    def reduce_max_sum(ms_list, cnt_list):
        if world == 1:
            return ms_list, cnt_list
        t = torch.tensor(ms_list, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor(cnt_list, dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        return [float(x) for x in t], [float(x) for x in c]

    KERNELS = {7: "brw_box_metropolis_byte_epoch_kernel (EXACT: reference association for every trial)",
               6: "brw_box_metropolis_byte_epoch_kernel",
               5: "brw_box_metropolis_{k}_kernel (EXACT: reference association for every trial)",
               4: "brw_box_metropolis_{k}_kernel", 3: "brw_box_metropolis_fast_kernel<screened>",
               2: "brw_box_metropolis_fast_kernel", 1: "brw_box_metropolis_kernel", 0: "brw_chain_metropolis_kernel"}

    def roofline_block(plan, per_launch_trials, per_launch_ms, layout, b_alg=B_ALG):
        peak, peak_src = peak_hbm()
        achieved = per_launch_trials * b_alg / (per_launch_ms * 1e-3) / 1e9
        epoch = plan["use_box"] in (4, 5) and plan["trials_per_step"] == 960
        prof = {}
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                prof = json.load(open(tp))
            except Exception:
                prof = {}
        kern = KERNELS[plan["use_box"]].format(k="epoch" if epoch else "word")
        if plan["use_box"] in (6, 7):
            # byte-lattice epoch kernels: the ncu capture of the matching geometry, if one is committed
            key = {(6, 6, 6): "bcc_6shell", (4, 4, 6): "fcc_4shell", (4, 6, 4): "fcc_4shell", (6, 4, 4): "fcc_4shell"}.get(tuple(plan["P"]))
            prof = prof.get("byte_epoch", {}).get(key, {}) if b_alg != 2 * 86 + 4 else {}
        elif not epoch:
            prof = {}
        return {
            # SURVEY 8(d) nominal figure: algorithmic bytes (2Z+4 per attempted swap) over the measured HBM copy peak.  The
            # lattice is L2- and shared-memory-resident by design, so HBM is NOT what binds these kernels (traffic below is
            # one read + one write of the lattice per launch); what does -- measured with ncu, profiles/ -- is named in `limiter`.
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "frac_hbm_nominal": achieved / peak,
            "traffic": prof.get("dram_bytes_per_launch"),
            "limiter": prof.get("limiter", "not profiled for this kernel / geometry"),
            "ncu": prof.get("ncu"),
            "kernel": kern,
            "algorithmic_bytes_per_attempt": b_alg, "attempts_per_launch": per_launch_trials,
            "ms_per_launch": per_launch_ms, "peak_source": peak_src}

    # ---- headline: device-resident arm ("value") ------------------------------------------------------
    workload = args.workload
    dev, beta, n, R, S, V, desc = make_device(workload, args.layout)
    if beta is None:
        beta = 1.0 / (T_KELVIN * brawl_b200.K_B_IN_RY)
    N = 2 * n ** 3
    plan = dev.metropolis_plan()
    e_start = dev.total_energy(0, R, exact_order=False).mean()
    trials_per_step = args.sweeps * N
    sampler = ClockSampler(local_rank)
    attempts, total_ms, launches, acceptance = timed(dev, beta, trials_per_step, args.steps, args.warmup, sampler=sampler)
    clocks = sampler.stop()
    attempts_rank = attempts * R                                 # planned attempts are per replica
    e_end = dev.total_energy(0, R, exact_order=False).mean()

    # ---- end-to-end arm ("e2e"): host buffers through the public C-ABI calls ----------------------
    # Every step: pinned host configuration -> set_config (H2D + pack), the trials, get_config (unpack + D2H) and the
    # total energy (the step's result).  Two lattices ping-pong: while lattice A's trials run on the compute stream, the
    # copy stream moves lattice B's result to the host and its next input to the device, so PCIe overlaps the kernels --
    # a Fortran driver annealing two replicas per rank would make exactly these calls.
    R2 = 2 if workload == "chain" else 1
    devs = [dev] if R2 == 1 else [dev, make_device(workload, args.layout)[0]]
    copy_stream = torch.cuda.Stream()
    if R2 == 2:
        devs[1].set_stream(copy_stream.cuda_stream)              # second lattice: its own stream => overlaps with the first
    host_cfg = [torch.empty((R, 2 * n, 2 * n, 2 * n), dtype=torch.int8).pin_memory() for _ in devs]
    host_np = [h.numpy() for h in host_cfg]
    host_lat = [torch.empty((R, N), dtype=torch.uint8).pin_memory() for _ in devs]
    host_lat_np = [h.numpy() for h in host_lat]
    for d, hnp, hl in zip(devs, host_np, host_lat_np):
        d.get_config(0, R, out=hnp)
        d.get_lattice(0, R, out=hl)
    import concurrent.futures as cf
    pool = cf.ThreadPoolExecutor(max_workers=len(devs))          # ctypes releases the GIL: one host thread per lattice

    def e2e_step(i, grid):
        d = devs[i]
        if grid:                                                 # the reference's `config` grid: 8 n^3 bytes each way
            d.set_config(host_np[i])                             # H2D + pack kernel
        else:                                                    # compact sites: 2 n^3 bytes each way
            d.set_lattice(host_lat_np[i])                        # H2D + validation kernel
        a, c, _ = d.metropolis_run(beta, trials_per_step)        # trials
        if grid:
            d.get_config(0, R, out=host_np[i])                   # unpack kernel + D2H
        else:
            d.get_lattice(0, R, out=host_lat_np[i])              # D2H
        e_host = d.total_energy(0, R, exact_order=False)         # D2H 8 bytes per replica (2 kernels)
        return int(a.sum()), d.metropolis_last_launches() + (4 if grid else 3), float(e_host[0])

    def e2e_run(grid):
        ms, att, nl = 0.0, 0, 0
        n_e2e = max(3, min(args.steps, 5))
        for it in range(2 + n_e2e):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = list(pool.map(lambda i: e2e_step(i, grid), range(len(devs))))
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) * 1e3                # host clock around device-synchronised calls (two streams)
            if it >= 2:
                ms += dt
                att += sum(r[0] for r in res)
                nl += sum(r[1] for r in res)
        return ms, att, nl

    barrier()
    e2e_ms, e2e_attempts, e2e_launches = e2e_run(False)
    barrier()
    e2e_grid_ms, e2e_grid_attempts, _ = e2e_run(True)
    pool.shutdown()

    (total_ms_r, e2e_ms_r, e2e_grid_ms_r), (attempts_all, e2e_attempts_all, launches_all, e2e_grid_attempts_all) = reduce_max_sum(
        [total_ms, e2e_ms, e2e_grid_ms], [attempts_rank, e2e_attempts, launches, e2e_grid_attempts])

    # ---- extra blocks: the other BASELINE configurations, short runs in the same clocks window ------------------
    extra = {}
    if not args.no_extra and workload == "chain":
        def short(name, wl_kind, layout, temp=None, nbr_swap=False, sweeps=32, steps=4, note=None):
            d, b, nn, RR, SS, VV, dsc = make_device(wl_kind, layout)
            if b is None:
                b = 1.0 / ((temp or T_KELVIN) * brawl_b200.K_B_IN_RY)
            pl = d.metropolis_plan(nbr_swap)
            tps = sweeps * d.n_atoms
            at, ms, nl, accr = timed(d, b, tps, steps, 3, nbr_swap=nbr_swap)
            (ms_r,), (at_all, nl_all) = reduce_max_sum([ms], [at * RR, nl])
            blk = {"metric": METRIC, "value": at_all / (ms_r * 1e-3), "unit": "swaps/s", "n_gpus": world,
                   "workload": dsc or ("AlTiCrMo bcc %d^3, 4 shells, T=%g K%s" % (nn, temp or T_KELVIN, ", nbr_swap=T" if nbr_swap else "")),
                   "acceptance": accr, "steps": steps, "sweeps_per_step": sweeps, "gpu_launches": int(nl_all),
                   "decomposition": pl,
                   "roofline": roofline_block(pl, at / max(1, nl), ms / max(1, nl), layout, getattr(d, "b_alg", B_ALG))}
            if note:
                blk["note"] = note
            extra[name] = blk
            d.close()
        short("T300", "chain", 0, temp=300.0)
        short("T3000", "chain", 0, temp=3000.0)
        short("nbr_swap", "chain", 0, nbr_swap=True, sweeps=8,
              note="nearest-neighbour swaps (metropolis.inp nbr_swap=T): generic byte-lattice box kernel")
        short("epoch8", "chain", 4,
              note="epochs of 8 steps: more attempts per second, but a site is tried 8 times against a frozen neighbourhood -- "
                   "sampling efficiency per attempt 0.27 of the sequential sampler vs 0.40-0.45 for the default (4 steps) and "
                   "for the round-1 kernel (tools/exp_scan.py, profiles/r02_sampling_efficiency.txt)")
        short("round1_kernel", "chain", 3, note="the round-1 default (one gather per step, two warp groups), same build")
        short("replicas", "replicas", 0, sweeps=16, steps=3)
        note6 = ("byte-lattice epoch kernel (period-P residue classes, site energies cached over 4 steps); sampling efficiency per "
                 "attempt ~0.3 of the sequential sampler (tools/exp_byte_epoch.py relaxation)")
        short("bcc_6shell", "bcc6", 0, note=note6)
        short("fcc_4shell_quinary", "fcc4", 0, note=note6)
        short("fcc_4shell_FeNi", "feni", 0, note=note6 + "; BASELINE configs[0] Hamiltonian at 128^3")
        short("fcc_6shell_quinary", "fcc6", 0, note=note6)
        dev.close()
        for d in devs[1:]:
            d.close()
        extra["wl"] = wl_time_to_flatness(rank, world, local_rank, args.wl_windows_per_gpu, args.wl_walkers)
        extra["ns"] = ns_walk_rate(rank, world, local_rank)

    if rank == 0:
        value = attempts_all / (total_ms_r * 1e-3)
        e2e_value = e2e_attempts_all / (e2e_ms_r * 1e-3)
        wl_name = desc or workload_name(n)
        out = {
            "metric": METRIC, "value": value, "unit": "swaps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_r / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name,
                       "attempted_swaps_per_step": trials_per_step * R, "sweeps_per_step": args.sweeps,
                       "l2": "flushed with a 512 MiB memset before every timed step",
                       "decomposition": plan,
                       "dE_mode": {0: "reference f64 association for every trial",
                                   1: "integer-count screening + reference association inside the guard band "
                                      "(accept/reject decisions identical to mode 0)",
                                   2: "dense non-interacting-set decomposition, site energies cached over epochs of 4 steps, "
                                      "fixed-point dE + f32 acceptance pre-test; inside its guard band (~1e-5 of the trials) f64 "
                                      "dE from the exact integer counts, inside that one's band (~1e-9) the reference f64 "
                                      "association (accept/reject decisions identical to mode 0)"}[args.dE_mode],
                       "acceptance": acceptance,
                       # same-species attempts count as attempts but touch 2 B and no flops
                       "distinct_species_fraction": 1.0 - 1.0 / S,
                       "sampling_efficiency_per_attempt_vs_sequential": "0.40-0.45 (energy relaxation at 1000 K against the oracle's "
                                                                        "sequential sampler; round-1 kernel: 0.45; tools/exp_scan.py)",
                       "energy_per_atom_start_end_Ry": [e_start / N, e_end / N]},
            "e2e": {"value": e2e_value, "unit": "swaps/s", "h2d_bytes_per_step": int(len(devs) * R * (N + 8)),
                    "d2h_bytes_per_step": int(len(devs) * R * (N + 8 + 24)),
                    "calls": "set_lattice + metropolis_run + get_lattice + total_energy per lattice through the C ABI, pinned "
                             "host buffers of compact sites (1 B per atom); %d lattice(s) per GPU on separate streams so "
                             "that copies overlap the trials" % len(devs),
                    "with_reference_config_grid": {
                        "value": e2e_grid_attempts_all / (e2e_grid_ms_r * 1e-3), "unit": "swaps/s",
                        "h2d_bytes_per_step": int(len(devs) * R * (8 * n ** 3 + 8)),
                        "d2h_bytes_per_step": int(len(devs) * R * (8 * n ** 3 + 8 + 24)),
                        "calls": "set_config + metropolis_run + get_config + total_energy: the reference's int8 config grid "
                                 "(2n)^3, 4x the bytes of the compact form, pack / unpack kernels on the device"}},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": roofline_block(plan, attempts / max(1, launches), total_ms / max(1, launches), args.layout),
        }
        if extra:
            out["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            if workload == "replicas":
                v, dt = cpu_reference_rate(cores, args.cpu_trials, n=n, S=S, V=V, temp=1000.0)
            else:
                v, dt = cpu_reference_rate(cores, args.cpu_trials)
            out["cpu_baseline"] = {"value": v, "unit": "swaps/s", "cores": cores, "kind": "port",
                                   "sample": "%d threads x %d trials on private %d^3 bcc replicas, %.1f s; C restatement "
                                             "of the reference hot path (bit-exact vs reference goldens)" % (cores, args.cpu_trials, n, dt)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
