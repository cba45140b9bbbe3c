"""Wang-Landau on the GPU: the reference's regression case tests/04_parallel_wang-landau
(bcc n=4, 4 species, 6 shells, 512 bins, 4 windows, overlap 0.25, f 0.05 -> 5e-5, flatness 0.9)
against its golden ln g(E) with the reference's own acceptance criterion (NRMSE < 1 %,
tests/ci_test.py:42-50)."""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def nrmse(ref, test):
    return float(np.sqrt(np.mean((ref - test) ** 2)) / np.mean(np.abs(ref)))


def test_enter_energy_window(orc, golden):
    import brawl_b200
    from brawl_b200 import wang_landau as wl
    p = wl.WLParams()
    drv = wl.WangLandau("bcc", 4, 4, 4, 4, 6, golden["t04_V"], [32, 32, 32, 32], p, walkers=4)
    drv.enter_energy_windows()
    e = drv.dev.total_energy(0, drv.n_local)
    assert np.allclose(e, drv.energies, rtol=0, atol=1e-11)          # running energy == exact energy
    lo = drv.edges[drv.win_lo - 1]; hi = drv.edges[drv.win_hi]
    cond = 0.1 * np.abs(hi - lo)
    assert np.all((e > lo + cond) & (e < hi - cond))
    g = drv.dev.get_config(0, drv.n_local)
    for gi in g:
        assert np.array_equal(np.bincount(gi.ravel(), minlength=5), [384, 32, 32, 32, 32])


def test_wang_landau_golden_04(orc, golden):
    import brawl_b200
    from brawl_b200 import wang_landau as wl
    p = wl.WLParams(mc_sweeps=100, bins=512, num_windows=4, bin_overlap=0.25, tolerance=5e-5, flatness=0.90,
                    wl_f=0.05, energy_min=-96, energy_max=0.0, radial_samples=8)
    drv = wl.WangLandau("bcc", 4, 4, 4, 4, 6, golden["t04_V"], [32, 32, 32, 32], p, walkers=8, seed=2024)
    t0 = time.time()
    lng = drv.run()
    dt = time.time() - t0
    ref = np.asarray(golden["t04_wl_dos"], dtype=np.float64)
    err = nrmse(ref, lng)
    print("WL: %.1f s, %d sweeps calls per stage %s, %.3g trials, NRMSE %.4f" % (dt, sum(drv.stage_sweeps), drv.stage_sweeps, drv.total_trials, err))
    assert lng.min() == 0.0 and np.all(np.isfinite(lng))
    assert err < 0.01, err


def test_wang_landau_dynamic_windows_golden_04(orc, golden):
    """The reference's default `performance = 0`: windows resized by mpi_window_optimise (wang-landau.F90:1211-1328)
    after pre-sampling and after every f-stage, walkers steered into the new windows on the GPU; same golden ln g(E),
    same 1 % NRMSE criterion (tests/ci_test.py:42-50)."""
    from brawl_b200 import wang_landau as wl
    p = wl.WLParams(mc_sweeps=100, bins=512, num_windows=4, bin_overlap=0.25, tolerance=5e-5, flatness=0.90,
                    wl_f=0.05, energy_min=-96, energy_max=0.0, radial_samples=8, performance=0)
    drv = wl.WangLandau("bcc", 4, 4, 4, 4, 6, golden["t04_V"], [32, 32, 32, 32], p, walkers=8, seed=2025)
    t0 = time.time()
    lng = drv.run()
    dt = time.time() - t0
    ref = np.asarray(golden["t04_wl_dos"], dtype=np.float64)
    err = nrmse(ref, lng)
    widths = [(h[:, 1] - h[:, 0] + 1).tolist() for h in drv.window_history]
    print("WL dynamic windows: %.1f s, sweeps per stage %s, NRMSE %.4f, widths %s -> %s" % (dt, drv.stage_sweeps, err, widths[0], widths[-1]))
    assert len(drv.window_history) == 1 + len(drv.stage_sweeps)
    for h in drv.window_history:
        assert h[0, 0] == 1 and h[-1, 1] == 512 and np.all(h[1:, 0] <= h[:-1, 1])
    e = drv.dev.total_energy(0, drv.n_local)
    lo = drv.edges[drv.win_lo - 1]; hi = drv.edges[drv.win_hi]
    assert np.all((e > lo) & (e < hi)) and np.allclose(e, drv.energies, rtol=0, atol=1e-11)
    assert lng.min() == 0.0 and np.all(np.isfinite(lng))
    assert err < 0.01, err
    # <E>(T) from the final ln g (compute_mean_energy, :457-477) against the same sum over the golden ln g
    me_ref = wl.compute_mean_energy(ref, drv.edges, 512, drv.bin_width)
    assert np.max(np.abs(drv.mean_energy[29:, 0] - me_ref[29:, 0])) < 0.02 * abs(drv.edges[0])      # T >= 300 K


def test_wang_landau_rho_of_E_on_gpu(orc, golden):
    """rho(E) sampled with the SRO kernel during a short WL run (wang-landau.F90:574-592, save_rho_E :346-381): sum
    rules per bin, agreement of a walker's sample with the oracle's radial_densities of the same configuration, and a
    first-shell like-pair density that changes between the low- and the high-energy bins."""
    from brawl_b200 import wang_landau as wl
    p = wl.WLParams(mc_sweeps=50, bins=64, num_windows=2, bin_overlap=0.25, tolerance=0.02, flatness=0.8, wl_f=0.05,
                    energy_min=-60.0, energy_max=0.0, radial_samples=8, performance=0)
    drv = wl.WangLandau("bcc", 4, 4, 4, 4, 6, golden["t04_V"], [32, 32, 32, 32], p, walkers=8, seed=11, wc_range=3)
    drv.run(max_sweeps_per_stage=2000)
    rho, n = drv.rho_of_E_partial()
    assert np.count_nonzero(n) > 40 and drv.radial_record.max() <= 1 and 0.0 < drv.radial_min <= 1.0
    for b in np.flatnonzero(n):
        for l, z in enumerate((1, 8, 6)):
            assert np.allclose(rho[b, l].sum(axis=0), z, atol=1e-9) or np.allclose(rho[b, l].sum(axis=1), z, atol=1e-9)
    sysm = orc.System("bcc", 4, 4, 4, 4, 6, golden["t04_V"])
    g = drv.dev.get_config(3)
    assert np.array_equal(drv.dev.radial_densities(3, 3), sysm.radial_densities(g, 3, sysm.lattice_shells(g, 3)))
    like = np.array([np.trace(rho[b, 1]) for b in np.flatnonzero(n)])       # like-pair density in the first shell
    assert np.all(np.isfinite(like)) and abs(like[:5].mean() - like[-5:].mean()) > 1e-3     # SRO changes with E


@pytest.mark.parametrize("n,lo_bin,hi_bin,max_iters", [(4, 200, 260, 60000), (4, 30, 60, 40000), (2, 1, 6, 12000)])
def test_enter_energy_window_replay_matches_oracle(orc, golden, n, lo_bin, hi_bin, max_iters):
    """enter_energy_window (wang-landau.F90:643-741) on the reference's MT19937 stream: the GPU replay kernel driven by the
    host loop the header prescribes (exact-energy re-check on status 1, initial_setup on status 2) against the oracle's
    restatement -- same exit, same iteration count, bit-identical running energy, configuration and MT state.  The
    n = 2 case (16 atoms: re-randomisation every 4000 iterations, window below the reachable energies) crosses the
    initial_setup branch with its stale running energy three times and stops at max_iters without entering."""
    import brawl_b200
    from brawl_b200 import wang_landau as wl
    V = golden["t04_V"]
    sysm = orc.System("bcc", n, n, n, 4, 6, V)
    N = sysm.n_atoms
    conc, cnt = sysm.quotas(conc=[0.25] * 4)
    edges = wl.create_energy_bins(N, -96.0, 0.0, 512)
    min_e, max_e = float(edges[lo_bin - 1]), float(edges[hi_bin])
    # oracle
    mt_o = orc.MT(rank=3)
    g_o = sysm.initial_setup(mt_o, conc, cnt)
    g0 = g_o.copy()
    st0 = mt_o.state625().copy()
    ok_o, e_o, it_o = sysm.wl_enter_energy_window(g_o, mt_o, conc, cnt, min_e, max_e, -96.0, 0.0, max_iters)
    # GPU replay + the host loop
    dev = brawl_b200.Device("bcc", n, n, n, 4, 6, V)
    dev.set_config(g0)
    mt_h = orc.MT(seed=1); mt_h.load625(st0)            # host copy of the stream for initial_setup (status 2)
    st = st0.copy()
    target, cond = (min_e + max_e) / 2.0, abs(max_e - min_e) * 0.1
    sf = np.float32(0.0025) * np.float32(abs(np.float32(0.0) - np.float32(-96.0)))
    sf = np.float32(sf * np.float32(N))
    sigma = float(sf) / (13.605693122 * 1000.0)
    e = float(dev.total_energy(exact_order=True)[0])
    i_steps, it, resume, entered, n_rerand = 0, 0, 0, False, 0
    while it < max_iters or resume:
        status, e, i_steps, done = dev.wl_enter_window_replay(e, target, min_e + cond, max_e - cond, 2.0 * (sigma * sigma),
                                                              N * 250, i_steps, max_iters - it, resume, st)
        it += done
        resume = 0
        if status == 1:
            e = float(dev.total_energy(exact_order=True)[0])                 # :692
            if max_e - cond > e > min_e + cond:
                entered = True
                break
        elif status == 2:                                                    # :677-680
            mt_h.load625(st)
            dev.set_config(sysm.initial_setup(mt_h, conc, cnt))
            st = mt_h.state625().copy()
            resume, n_rerand = 1, n_rerand + 1
        else:
            break
    assert (entered, it) == (ok_o, it_o), (entered, it, ok_o, it_o)
    assert e == e_o
    assert np.array_equal(dev.get_config(), g_o)
    assert np.array_equal(st, mt_o.state625())
    if n == 2:
        assert n_rerand >= 2 and not entered


def test_wl_set_span_needs_a_communicator(golden):
    """brawl_cuda_wl_set_span: windows shared by several ranks need the handle's NCCL communicator; one rank is the default."""
    import brawl_b200 as bw
    dev = bw.Device("bcc", 4, 4, 4, 4, 6, golden["t04_V"], n_replicas=4)
    with pytest.raises(bw.BrawlCudaError, match="wl_init"):
        dev.wl_set_span(2)
    dev.wl_init(64, np.linspace(-1.0, 0.0, 65), 2)
    dev.wl_set_span(1)
    with pytest.raises(bw.BrawlCudaError, match="communicator"):
        dev.wl_set_span(2)
    dev.close()


def test_wang_landau_windows_spanning_two_gpus(golden):
    """Three windows on two GPUs (not shardable): every GPU holds 8 walkers of every window and the window average of each
    `sweeps` call is an ncclAllReduce over the ABI's communicator (brawl_cuda_wl_set_span).  Reference case 04 within its
    own criterion.  Needs two devices (gpurun --gpus 2)."""
    import ctypes as C
    import json
    import subprocess
    import sys
    n = C.c_int(0)
    lib = C.CDLL(os.path.join(ROOT, "brawl_b200", "libbrawl_cuda.so"))
    if lib.brawl_cuda_device_count(C.byref(n)) != 0 or n.value < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "wl_multi_gpu.py"), "--windows", "3", "--walkers", "8",
                        "--span"], capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and line, r.stderr[-2000:]
    out = json.loads(line[-1])
    assert out["span"] and out["n_gpus"] == 2 and out["walkers_per_window"] == 16 and out["comm"] == "abi"
    assert out["nrmse_vs_reference_golden"] < 0.01
