"""CPU tests of the Wang-Landau host logic (window division, stitching, exchange planning) incl. a
world_size-2 gloo run of the collective plumbing."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_window_division_matches_reference_arithmetic():
    from brawl_b200 import wang_landau as wl
    iv = wl.divide_range(512, 4)                       # wang-landau.F90:855-875
    assert iv.tolist() == [[1, 128], [129, 256], [257, 384], [385, 512]]
    idx = wl.create_overlap(iv, np.float32(0.25))      # :934-955 (note the last window's width quirk)
    assert idx.tolist() == [[1, 128], [97, 256], [217, 384], [343, 512]]
    idx0 = wl.create_overlap(iv, np.float32(0.0))      # overlap 0 still extends windows by 2 bins
    assert idx0.tolist() == [[1, 128], [127, 256], [255, 384], [383, 512]]
    assert wl.divide_range(512, 1).tolist() == [[1, 512]]
    p = wl.WLParams(wl_f=0.05, tolerance=5e-5)
    assert p.wl_f == 0.05000000074505806               # single-precision wl_params


def test_energy_bins_and_bin_index(orc, golden):
    from brawl_b200 import wang_landau as wl
    sysm = orc.System("bcc", 4, 4, 4, 4, 6, golden["t04_V"])
    e_ref = sysm.wl_bin_edges(-96.0, 0.0, 512)
    e = wl.create_energy_bins(128, -96.0, 0.0, 512)
    assert np.array_equal(e, e_ref)
    for x in np.linspace(e[0] * 1.01, e[-1] + 1e-4, 300):
        assert wl.bin_index(x, e, 512) == orc.lib().orc_bin_index(__import__("ctypes").c_double(x), e.ctypes.data_as(__import__("ctypes").c_void_p), 512)


def test_dos_combine_recovers_a_smooth_curve():
    from brawl_b200 import wang_landau as wl
    bins = 512
    x = np.linspace(0, 1, bins)
    true = 140 * np.sqrt(x + 1e-3) - 20 * x ** 2
    idx = wl.create_overlap(wl.divide_range(bins, 4), np.float32(0.25))
    lw = np.zeros((4, bins))
    for q in range(4):
        lo, hi = idx[q]
        lw[q, lo - 1:hi] = true[lo - 1:hi] + 37.0 * q - 5.0      # arbitrary per-window offsets
    comb = wl.dos_combine(lw, idx)
    assert np.allclose(comb, true - true.min(), atol=1e-9)


def test_exchange_plan_is_deterministic_and_local():
    from brawl_b200 import wang_landau as wl
    bins, W, w = 512, 4, 3
    idx = wl.create_overlap(wl.divide_range(bins, W), np.float32(0.25))
    edges = wl.create_energy_bins(128, -96.0, 0.0, bins)
    centre = lambda b: 0.5 * (edges[b - 1] + edges[b])
    # window 1 walkers in bins (100,110,50), window 2 walkers in (100,230,240), window 3 (220,300,350), window 4 (345,400,500)
    e = [centre(b) for b in (100, 110, 50, 100, 230, 240, 220, 300, 350, 345, 400, 500)]
    lng = np.zeros((W, bins))
    plans = [wl.plan_replica_exchange(list(e), lng, idx, w, edges, np.random.default_rng(5)) for _ in range(2)]
    assert plans[0] == plans[1]
    for a, b in plans[0]:
        assert b // w == a // w + 1                    # adjacent windows only
    pairs = {(a // w, b // w) for a, b in plans[0]}
    assert (0, 1) in pairs and (1, 2) in pairs and (2, 3) in pairs    # flat ln g: every eligible pair accepted
    used = [x for ab in plans[0] for x in ab]
    assert len(used) == len(set(used))


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from brawl_b200 import wang_landau as wl
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    comm = wl._Comm(rank, world, None)
    bins, W, w = 512, 4, 2
    idx = wl.create_overlap(wl.divide_range(bins, W), np.float32(0.25))
    edges = wl.create_energy_bins(128, -96.0, 0.0, bins)
    rng_local = np.random.default_rng(100 + rank)
    # each rank owns 2 windows x 2 walkers; energies placed in this rank's windows
    my = []
    for q_ in range(rank * 2, rank * 2 + 2):
        lo, hi = idx[q_]
        my += [0.5 * (edges[b - 1] + edges[b]) for b in rng_local.integers(lo, hi + 1, size=w)]
    e_all = comm.all_gather(np.array(my)).reshape(-1)
    lng_local = np.full((2, bins), float(rank))
    lng_all = comm.all_gather(lng_local).reshape(W, bins)
    plan = wl.plan_replica_exchange(list(e_all), lng_all, idx, w, edges, np.random.default_rng(9))
    tot = comm.all_sum(rank + 1)
    q.put((rank, e_all.tolist(), plan, tot, lng_all[:, 0].tolist()))
    dist.destroy_process_group()


def test_collectives_world_size_2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] and len(res[0][1]) == 8      # same gathered energies on both ranks
    assert res[0][2] == res[1][2]                               # same exchange plan on both ranks
    assert res[0][3] == res[1][3] == 3.0
    assert res[0][4] == [0.0, 0.0, 1.0, 1.0]
