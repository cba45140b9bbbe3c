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


def test_window_optimise_matches_oracle_and_keeps_invariants(orc):
    """mpi_window_optimise arithmetic (wang-landau.F90:1224-1311): the host mirror against the oracle's C restatement
    on seeded inputs, plus the invariants the reference relies on (contiguous cover of 1..bins, minimum width)."""
    from brawl_b200 import wang_landau as wl
    rng = np.random.default_rng(11)
    for case in range(300):
        W = int(rng.integers(2, 17))
        bins = int(rng.choice([64, 128, 512, 1000]))
        if max(int(0.02 * bins), 2) * W > bins:
            continue
        iv = wl.divide_range(bins, W)
        prev = np.full(W, 1.0 / np.float32(W))
        spread = 10.0 ** rng.uniform(0, 3)
        for it in range(0, 4):
            mc = np.floor(rng.uniform(1, spread, W)) * 100 * 128
            a_iv, a_prev = wl.window_optimise(it, iv, mc, prev, bins)
            b_iv, b_prev = orc.wl_window_optimise(it, iv, mc, prev, bins)
            assert np.array_equal(a_iv, b_iv), (case, it)
            assert np.array_equal(a_prev, b_prev)
            assert a_iv[0, 0] == 1 and a_iv[-1, 1] == bins
            assert np.all(a_iv[1:, 0] == a_iv[:-1, 1] + 1)
            widths = a_iv[:, 1] - a_iv[:, 0] + 1
            assert widths.sum() == bins and widths.min() >= max(int(0.02 * bins), 2)
            assert abs(a_prev.sum() - 1.0) < 1e-12
            iv, prev = a_iv, a_prev
    # equal effort and equal windows: nothing moves (alpha = 1 after pre-sampling)
    iv = wl.divide_range(512, 4)
    out, prev = wl.window_optimise(0, iv, [1e6] * 4, [0.25] * 4, 512)
    assert out.tolist() == iv.tolist() and np.allclose(prev, 0.25)
    # a window that needed 3x the trials per bin shrinks, the others grow; later iterations damp the change
    out, prev = wl.window_optimise(0, iv, [3e6, 1e6, 1e6, 1e6], [0.25] * 4, 512)
    w = out[:, 1] - out[:, 0] + 1
    assert w[0] < 128 and np.all(w[1:] > 128) and w.sum() == 512
    out2, _ = wl.window_optimise(3, iv, [3e6, 1e6, 1e6, 1e6], [0.25] * 4, 512)
    w2 = out2[:, 1] - out2[:, 0] + 1
    assert w[0] < w2[0] < 128
    assert wl.sort_descending([5, 9, 5, 9, 1]) == [2, 4, 3, 1, 5]          # hand-traced exchange sort: not stable
    assert wl.window_optimise(0, [[1, 512]], [1e6], [1.0], 512)[0].tolist() == [[1, 512]]      # one window: untouched


def test_compute_mean_energy_matches_oracle(orc, golden):
    """compute_mean_energy (wang-landau.F90:457-477) on the reference's own ln g(E) of regression case 04."""
    from brawl_b200 import wang_landau as wl
    lng = np.asarray(golden["t04_wl_dos"], dtype=np.float64)
    edges = wl.create_energy_bins(128, -96.0, 0.0, 512)
    width = wl.energy_bin_width(128, -96.0, 0.0, 512)
    me = wl.compute_mean_energy(lng, edges, 512, width)
    ref = orc.wl_mean_energy(lng, edges, 512, width)
    assert me.shape == (300, 2) and np.array_equal(me[:, 1], ref[:, 1])
    assert np.allclose(me[:, 0], ref[:, 0], rtol=1e-13, atol=0)         # libm exp vs numpy exp: last-ulp differences only
    assert np.all(np.diff(me[:, 0]) > 0)                                 # <E>(T) rises with T
    assert edges[0] < me[0, 0] < me[-1, 0] < edges[-1]
    assert me[0, 1] == 1.0 / (wl.K_B_IN_RY * 1 * 10.0)


class _OracleDevice:
    """Stand-in for brawl_b200.Device in the CPU test of the WL driver's control flow: the same host-facing
    methods, with the trial loops run by the oracle (tests may use the oracle; the product never does)."""
    orc = None

    def __init__(self, lattice, n_1, n_2, n_3, n_species, n_shells, V_ex, device=0, n_replicas=1):
        o = self.orc
        self.sys = o.System(lattice, n_1, n_2, n_3, n_species, n_shells, V_ex)
        self.n_atoms, self.n_replicas = self.sys.n_atoms, n_replicas
        # one buffer for all replicas; self.g[r] are views, so lattice_tensor() can expose it to torch.distributed
        self.buf = np.zeros((n_replicas, 2 * n_3, 2 * n_2, 2 * n_1), dtype=np.int8)
        self.g = [self.buf[r] for r in range(n_replicas)]
        self.mt = [o.MT(seed=1000 + r + 17 * device) for r in range(n_replicas)]
        self.rng = np.random.default_rng(77 + device)

    def set_config(self, config, first_replica=0, n=None):
        self.g[first_replica][...] = np.asarray(config, dtype=np.int8).reshape(self.g[first_replica].shape)

    def lattice_tensor(self, torch):
        return torch.from_numpy(self.buf.reshape(self.n_replicas, -1))

    def random_config(self, species_count, first_replica=0, n=1, seed=0, offset=0):
        from brawl_b200 import wang_landau as wl
        for r in range(first_replica, first_replica + n):
            self.g[r][...] = wl.random_configuration("bcc", 4, 4, 4, species_count, self.rng)

    def radial_densities(self, wc_range, replica=0):
        g = self.g[replica]
        return self.sys.radial_densities(g, wc_range, self.sys.lattice_shells(g, wc_range))

    def swap_replicas(self, a, b):
        t = self.g[a].copy()
        self.g[a][...] = self.g[b]
        self.g[b][...] = t

    # device-resident Wang-Landau state of the product handle (brawl_cuda_wl_init / wl_iterate), kept in numpy here
    def wl_init(self, bins, bin_edges, walkers_per_window):
        n = len(self.g)
        self.bins, self.wpw, self.edges = bins, walkers_per_window, np.asarray(bin_edges, dtype=np.float64)
        self.lngs, self.hists = np.zeros((n, bins)), np.zeros((n, bins))
        self.lo, self.hi = np.ones(n, dtype=np.int32), np.full(n, bins, dtype=np.int32)

    def wl_set_windows(self, win_lo, win_hi, zero_hist=True):
        self.lo, self.hi = np.array(win_lo, dtype=np.int32), np.array(win_hi, dtype=np.int32)
        if zero_hist:
            self.hists[...] = 0.0

    def wl_zero_hist(self):
        self.hists[...] = 0.0

    def wl_set_span(self, n_ranks):
        self.span = int(n_ranks)                 # brawl_cuda_wl_set_span: the windows are shared by n_ranks ranks

    def wl_set_lng(self, lng):
        self.lngs[...] = np.asarray(lng)[None, :]

    def wl_get(self, what=0):
        return (self.hists if what else self.lngs)[::self.wpw].copy()

    def synchronize(self):
        pass

    def wl_iterate(self, wl_f, n_trials, seed=0, offset=0, nbr_swap=False):
        n = len(self.g)
        ef = np.zeros(n)
        for w in range(n):
            nb = int(self.hi[w] - self.lo[w] + 1)
            h = np.ascontiguousarray(self.hists[w, :nb])
            _, ef[w] = self.sys.wl_sweeps(self.g[w], self.mt[w], self.lngs[w], h, self.edges, self.lo[w], self.hi[w],
                                          wl_f, n_trials, nbr_swap)
            self.hists[w, :nb] = h
        nq = n // self.wpw
        span = getattr(self, "span", 1)
        mn, mean = np.zeros(nq), np.zeros(nq)
        for q in range(nq):                      # the intra-window average (:628-631): local sums / total walkers ...
            sl = slice(q * self.wpw, (q + 1) * self.wpw)
            self.lngs[sl] = self.lngs[sl].sum(axis=0) / float(np.float32(self.wpw * span))
            self.hists[sl] = self.hists[sl].sum(axis=0) / float(np.float32(self.wpw * span))
        if span > 1:                             # ... summed over the ranks that share the windows (the ABI: ncclAllReduce)
            import torch
            import torch.distributed as dist
            for a in (self.lngs, self.hists):
                t = torch.from_numpy(a)
                dist.all_reduce(t)
        for q in range(nq):                      # the flatness inputs (:222-226)
            nb = int(self.hi[q * self.wpw] - self.lo[q * self.wpw] + 1)
            mn[q], mean[q] = self.hists[q * self.wpw, :nb].min(), self.hists[q * self.wpw, :nb].sum() / nb
        return ef, mn, mean

    def wl_enter_window(self, target, lo_e, hi_e, inv_two_sigma_sq, max_trials, seed=0, offset=0):
        e_out, ent = np.zeros(len(self.g)), np.zeros(len(self.g), dtype=np.int32)
        for w, g in enumerate(self.g):
            sites = np.argwhere(g > 0)
            e = self.sys.total_energy(g)
            for _ in range(int(max_trials)):
                if lo_e[w] < e < hi_e[w]:
                    break
                a, b = (tuple(s) for s in sites[self.rng.integers(0, len(sites), 2)])
                if g[a] == g[b]:
                    continue
                g[a], g[b] = g[b], g[a]
                e2 = self.sys.total_energy(g)
                if np.log(self.rng.random()) < -((e2 - target[w]) ** 2 - (e - target[w]) ** 2) * inv_two_sigma_sq:
                    e = e2
                else:
                    g[a], g[b] = g[b], g[a]
            e_out[w], ent[w] = e, int(lo_e[w] < e < hi_e[w])
        return e_out, ent


def test_driver_dynamic_windows_control_flow(orc, golden, monkeypatch, tmp_path):
    """performance = 0: windows are resized after pre-sampling and after every f-stage (wang-landau.F90:284-286, 820),
    walkers end every stage inside their (new) window, ln g stays a sane stitched curve.  Trial loops by the oracle."""
    from brawl_b200 import wang_landau as wl
    _OracleDevice.orc = orc
    monkeypatch.setattr(wl, "Device", _OracleDevice)
    p = wl.WLParams(mc_sweeps=20, bins=64, num_windows=3, bin_overlap=0.25, tolerance=0.02, flatness=0.7, wl_f=0.05,
                    energy_min=-60.0, energy_max=-5.0, performance=0)
    drv = wl.WangLandau("bcc", 4, 4, 4, 4, 6, golden["t04_V"], [32, 32, 32, 32], p, walkers=2, seed=5)
    seen = []
    lng = drv.run(max_sweeps_per_stage=400, callback=lambda f, c: seen.append(drv.window_indices.copy()))
    assert len(drv.stage_sweeps) == 3 and len(seen) == 2                   # pre-sampling + f = 0.05, 0.025
    assert len(drv.window_history) == 1 + 3                               # one resize per stage
    for idx in drv.window_history:
        assert idx[0, 0] == 1 and idx[-1, 1] == 64
        assert np.all(idx[1:, 0] <= idx[:-1, 1])                          # neighbours overlap
    assert any(not np.array_equal(drv.window_history[0], h) for h in drv.window_history[1:])   # and they did move
    assert np.all(drv.last_mc_steps > 0) and np.all(drv.wl_mc_steps == 0)
    assert abs(drv.diffusion_prev.sum() - 1.0) < 1e-12
    lo, hi = drv.edges[drv.win_lo - 1], drv.edges[drv.win_hi]
    e = np.array([drv.dev.sys.total_energy(g) for g in drv.dev.g])
    assert np.all((e > lo) & (e < hi))                                    # every walker inside its resized window
    assert np.array_equal(e, drv.energies)
    assert lng.min() == 0.0 and np.all(np.isfinite(lng)) and np.all(np.diff(lng[8:40]) > -1.0)
    assert drv.mean_energy.shape == (300, 2) and np.all(np.diff(drv.mean_energy[:, 0]) >= 0)
    # output files of save_wl_data (wang-landau.F90:409-417)
    from scipy.io import netcdf_file
    drv.save_wl_data(str(tmp_path), lng)
    rd = lambda f: np.array(netcdf_file(str(tmp_path / "data" / f), "r", mmap=False).variables["grid data"].data)
    assert np.array_equal(rd("wl_dos.nc"), lng) and np.array_equal(rd("wl_dos_bins.nc"), drv.edges) and rd("wl_hist.nc").shape == (64,)
    # performance = 4: static windows
    p4 = wl.WLParams(mc_sweeps=20, bins=64, num_windows=3, bin_overlap=0.25, tolerance=0.04, flatness=0.7, wl_f=0.05,
                     energy_min=-60.0, energy_max=-5.0, performance=4)
    d4 = wl.WangLandau("bcc", 4, 4, 4, 4, 6, golden["t04_V"], [32, 32, 32, 32], p4, walkers=2, seed=5)
    d4.run(max_sweeps_per_stage=400)
    assert len(d4.window_history) == 1


def test_driver_rho_of_E_sampling(orc, golden, monkeypatch, tmp_path):
    """rho(E) (wang-landau.F90:574-592, save_rho_E :346-381): per-bin means of the radial densities, capped at
    max(radial_samples / walkers, 1) samples per walker and bin; a bin completes at radial_samples samples."""
    from brawl_b200 import wang_landau as wl
    _OracleDevice.orc = orc
    monkeypatch.setattr(wl, "Device", _OracleDevice)
    p = wl.WLParams(mc_sweeps=20, bins=32, num_windows=2, bin_overlap=0.25, tolerance=0.02, flatness=0.7, wl_f=0.05,
                    energy_min=-50.0, energy_max=-5.0, radial_samples=4, performance=4)
    drv = wl.WangLandau("bcc", 4, 4, 4, 4, 6, golden["t04_V"], [32, 32, 32, 32], p, walkers=2, seed=5, wc_range=3)
    drv.run(max_sweeps_per_stage=300)
    rho, n = drv.rho_of_E_partial()
    assert rho.shape == (32, 3, 4, 4) and n.shape == (32,)
    assert drv.radial_record.max() <= 2                                    # cap = max(4 // 2, 1) per walker and bin
    assert n.max() <= 2 * 2 * 2                                            # <= cap x walkers x 2 overlapping windows
    assert np.count_nonzero(n) > 16 and 0.0 < drv.radial_min <= 1.0
    assert np.array_equal(drv.radial_record_bool, n >= 4) or drv.rho_saved
    z = [1, 8, 6]             # lattice_shells starts at distance 0 (the atom itself, analytics.f90:205-275), then bcc 8, 6
    for b in np.flatnonzero(n):
        for l in range(3):
            # every atom has z_l neighbours in shell l: sum over the neighbour species of rho(i, j, l) = z_l
            assert np.allclose(rho[b, l].sum(axis=0), z[l]) or np.allclose(rho[b, l].sum(axis=1), z[l])
    # ordering tendency: the unlike-pair density in shell 1 differs between the lowest and highest sampled bins
    lo_b, hi_b = np.flatnonzero(n)[0], np.flatnonzero(n)[-1]
    assert not np.allclose(rho[lo_b, 1], rho[hi_b, 1])
    # asro/rho_of_E.nc (ncdf_radial_density_writer_across_energy, netcdf_io.f90:260-346)
    from scipy.io import netcdf_file
    setup = dict(n_1=4, n_2=4, n_3=4, n_species=4, lattice="bcc", interaction_file="bcc_epi.vij",
                 species_concentrations=[0.0, 0.25, 0.25, 0.25, 0.25], wc_range=3)
    shells = drv.dev.sys.lattice_shells(drv.dev.g[0], 3)
    drv.save_rho_of_E(str(tmp_path), shells, setup)
    f = netcdf_file(str(tmp_path / "asro" / "rho_of_E.nc"), "r", mmap=False)
    assert f.variables["rho data"].dimensions == ("U", "r", "j", "i") and np.array_equal(f.variables["rho data"].data, rho)
    assert np.array_equal(f.variables["r data"].data, shells)
    u = np.array(f.variables["U data"].data)
    assert u.shape == (32,) and np.allclose(u, 0.5 * (drv.edges[:-1] + drv.edges[1:]), rtol=1e-13)
    # off by default
    d0 = wl.WangLandau("bcc", 4, 4, 4, 4, 6, golden["t04_V"], [32, 32, 32, 32], p, walkers=2, seed=5)
    d0.run(max_sweeps_per_stage=50)
    assert d0.radial_record.sum() == 0 and d0.rho_of_E is None


def _gloo_driver_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import oracle
    from brawl_b200 import wang_landau as wl
    import test_wl_host_logic as t
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    t._OracleDevice.orc = oracle
    wl.Device = t._OracleDevice
    golden = np.load(os.path.join(ROOT, "tests", "golden", "brawl_golden.npz"))
    p = wl.WLParams(mc_sweeps=20, bins=64, num_windows=4, bin_overlap=0.25, tolerance=0.02, flatness=0.7, wl_f=0.05,
                    energy_min=-60.0, energy_max=-5.0, performance=0)
    drv = wl.WangLandau("bcc", 4, 4, 4, 4, 6, golden["t04_V"], [32, 32, 32, 32], p, walkers=2, device=rank, rank=rank,
                        world=world, seed=5)
    swaps = []
    orig = drv._replica_exchange
    drv._replica_exchange = lambda *a: swaps.append(orig(*a)) or swaps[-1]
    lng = drv.run(max_sweeps_per_stage=400)
    e = np.array([drv.dev.sys.total_energy(g) for g in drv.dev.g])
    lo, hi = drv.edges[drv.win_lo - 1], drv.edges[drv.win_hi]
    q.put((rank, [h.tolist() for h in drv.window_history], lng.tolist(), bool(np.array_equal(e, drv.energies)),
           bool(np.all((e > lo) & (e < hi))), int(sum(swaps)), drv.last_mc_steps.tolist(), drv.stage_sweeps))
    dist.destroy_process_group()


def test_driver_dynamic_windows_world_size_2_gloo():
    """Four windows sharded over two ranks (two each), performance = 0: both ranks derive the same resized windows from
    the all-gathered per-window trial counts, stitch the same ln g, exchange configurations across the rank boundary
    (isend/irecv of the lattice bytes) with consistent energy bookkeeping, and end with every walker in its window."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_driver_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=240) for _ in ps])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b = res
    assert a[1] == b[1] and len(a[1]) == 4 and any(h != a[1][0] for h in a[1][1:])       # same windows, and they moved
    assert a[2] == b[2] and min(a[2]) == 0.0                                              # same stitched ln g
    assert a[3] and b[3] and a[4] and b[4]                                                # energies == configs; inside windows
    assert a[5] == b[5] and a[5] > 0                                                      # same exchange plan, some accepted
    assert a[6] == b[6] and min(a[6]) > 0 and a[7] == b[7]


def _gloo_span_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import oracle
    from brawl_b200 import wang_landau as wl
    import test_wl_host_logic as t
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    t._OracleDevice.orc = oracle
    wl.Device = t._OracleDevice
    golden = np.load(os.path.join(ROOT, "tests", "golden", "brawl_golden.npz"))
    # three windows on two ranks: not shardable; every rank holds 2 walkers of every window (4 per window in total)
    p = wl.WLParams(mc_sweeps=20, bins=64, num_windows=3, bin_overlap=0.25, tolerance=0.02, flatness=0.7, wl_f=0.05,
                    energy_min=-60.0, energy_max=-5.0, performance=4)
    drv = wl.WangLandau("bcc", 4, 4, 4, 4, 6, golden["t04_V"], [32, 32, 32, 32], p, walkers=2, device=rank, rank=rank,
                        world=world, seed=5, span=True)
    swaps = []
    orig = drv._do_swaps
    drv._do_swaps = lambda sw, e: swaps.extend(sw) or orig(sw, e)
    lng = drv.run(max_sweeps_per_stage=400)
    e = np.array([drv.dev.sys.total_energy(g) for g in drv.dev.g])
    lo, hi = drv.edges[drv.win_lo - 1], drv.edges[drv.win_hi]
    # (running energies of the sweeps: equal to the exact total energies up to f64 rounding)
    q.put((rank, lng.tolist(), bool(np.allclose(e, drv.energies, rtol=0, atol=1e-12)), bool(np.all((e > lo) & (e < hi))), swaps,
           drv.last_mc_steps.tolist(), drv.stage_sweeps, drv.dev.wl_get(0).tolist(), drv.walkers_total, drv.n_local))
    dist.destroy_process_group()


def test_driver_windows_spanning_ranks_world_size_2_gloo():
    """span = True: three windows on two ranks (not shardable) -- every rank holds walkers of every window, the window
    average of a `sweeps` call is an all-reduce over the ranks (brawl_cuda_wl_set_span on the GPU), the exchange plan runs
    over the window-major list of all walkers incl. pairs on different ranks.  Both ranks end with the same ln g tables,
    the same stitched curve, the same plan, consistent energies, every walker inside its window."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_span_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=240) for _ in ps])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b = res
    assert a[1] == b[1] and min(a[1]) == 0.0 and max(a[1]) > 0.0                          # same stitched ln g
    assert a[2] and b[2] and a[3] and b[3]                                                # energies == configs; inside windows
    assert a[4] == b[4] and len(a[4]) > 0                                                 # same exchange plan, some accepted
    n_local = a[9]
    assert any(x // n_local != y // n_local for x, y in a[4])                             # ... incl. pairs on different ranks
    assert a[5] == b[5] and min(a[5]) > 0 and a[6] == b[6]                                # trial counts summed over the ranks
    assert a[7] == b[7] and a[8] == 4                                                     # all-reduced window tables agree


def test_ncdf_writer_1d_is_byte_identical_to_the_reference_file(golden, tmp_path):
    """ncdf_writer_1d (netcdf_io.f90:731-806): the reference's own tests/99_ref/04_parallel_wang-landau/wl_dos.nc,
    rewritten from its data by the host mirror, byte for byte (header and payload)."""
    from scipy.io import netcdf_file
    from brawl_b200 import wang_landau as wl
    path = str(tmp_path / "wl_dos.nc")
    wl.ncdf_writer_1d(path, golden["t04_wl_dos"])
    assert open(path, "rb").read() == golden["raw_t04_wl_dos_nc"].tobytes()
    f = netcdf_file(path, "r", mmap=False)
    assert list(f.dimensions.items()) == [("x", 512)] and np.array_equal(f.variables["grid data"].data, golden["t04_wl_dos"])
    wl.ncdf_writer_1d(path, np.arange(5.0))                              # another length: still a valid classic file
    assert np.array_equal(netcdf_file(path, "r", mmap=False).variables["grid data"].data, np.arange(5.0))


def test_dos_combine_matches_oracle_bit_for_bit(orc):
    """dos_combine (wang-landau.F90:1147-1194): the host mirror against the oracle's line-by-line restatement on random
    piecewise ln g with arbitrary per-window offsets and noise (so that the closest-slope bin is a real choice):
    identical stitch bins and identical f64 results, incl. the in-place rewrite of comb(beta_index)."""
    from brawl_b200 import wang_landau as wl
    rng = np.random.default_rng(7)
    for _ in range(200):
        W, bins = int(rng.integers(2, 9)), int(rng.choice([64, 128, 512]))
        win = wl.create_overlap(wl.divide_range(bins, W), float(rng.choice([0.0, 0.1, 0.25, 0.5])))
        base = np.cumsum(rng.normal(0.3, 1.0, bins))
        lng = np.zeros((W, bins))
        for q in range(W):
            lo, hi = win[q]
            lng[q, lo - 1:hi] = base[lo - 1:hi] + rng.normal(0, 0.05, hi - lo + 1) + rng.normal(0, 30)
        assert np.array_equal(wl.dos_combine(lng, win), orc.wl_dos_combine(lng, win))
