"""The C++ host side (brawl_b200/host): the reference's own input files in, the reference's own
output files out.  CPU part: parsers, MT19937, initial_setup, lattice_shells and the NetCDF-3 classic
writer (dryrun=1, no GPU) against the goldens.  GPU part: brawl_driver reproduces the reference's
regression cases 01, 02 (4 emulated MPI ranks) and 03 file for file."""
import os
import subprocess

import numpy as np
import pytest
from scipy.io import netcdf_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "brawl_b200", "host", "brawl_driver")


@pytest.fixture(scope="module")
def driver():
    from brawl_b200 import build
    build.build_library()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "brawl_b200", "host")])
    return DRIVER


def write_case(tmp, golden, case, files):
    for fn in files:
        with open(os.path.join(tmp, fn), "w") as fh:
            fh.write(str(golden["in_%s_%s" % (case, fn)]))


def nc_var(path, name):
    return np.array(netcdf_file(path, "r", mmap=False).variables[name].data)


def test_dryrun_parsers_and_initial_config(driver, golden, tmp_path):
    tmp = str(tmp_path)
    write_case(tmp, golden, "02", ("brawl.inp", "metropolis.inp", "bcc_epi.vij"))
    out = subprocess.run([driver, "dryrun=1"], cwd=tmp, capture_output=True, text=True, check=True).stdout
    assert "mode=301 lattice=bcc n=4,4,4 n_species=4 n_atoms=128 interaction_file=bcc_epi.vij interaction_range=6 wc_range=2 static_seed=1" in out
    assert "metropolis mode=simulated_annealing n_mc_steps=128 n_sample_steps=1 asro=128 alro=128 traj=1 T=300.000000 T_steps=1" in out
    V = golden["t02_V"]
    assert "V_ex entries=96 first=%.17g last=%.17g" % (V[0], V[95]) in out
    assert "shells 0 %.17g" % float(np.float32(np.sqrt(3.0))) in out
    # initial configuration: byte-identical NetCDF file to the reference's golden (header + data)
    mine = open(os.path.join(tmp, "configs", "dryrun_initial_config.nc"), "rb").read()
    assert mine == golden["raw_t02_r0_initial_nc"].tobytes()
    # fcc case with species_numbers (quota path of initialise.F90:468-472)
    tmp2 = str(tmp_path / "c01"); os.makedirs(tmp2)
    write_case(tmp2, golden, "01", ("brawl.inp", "metropolis.inp", "fcc_epi.vij"))
    subprocess.run([driver, "dryrun=1"], cwd=tmp2, capture_output=True, text=True, check=True)
    cfg = nc_var(os.path.join(tmp2, "configs", "dryrun_initial_config.nc"), "configuration")[..., 0]
    assert np.array_equal(cfg, golden["t01_initial"])


def test_driver_error_messages_match_reference(driver, golden, tmp_path):
    tmp = str(tmp_path)
    r = subprocess.run([driver, "dryrun=1"], cwd=tmp, capture_output=True, text=True)
    assert r.returncode != 0 and "Could not find input file brawl.inp" in r.stderr
    txt = str(golden["in_02_brawl.inp"]).replace("interaction_range = 6", "interaction_range = 11")
    open(os.path.join(tmp, "brawl.inp"), "w").write(txt)
    r = subprocess.run([driver, "dryrun=1"], cwd=tmp, capture_output=True, text=True)
    assert r.returncode != 0 and "Unsupported number of shells" in r.stderr
    txt = "\n".join(l for l in str(golden["in_02_brawl.inp"]).split("\n") if not l.startswith("n_species"))
    open(os.path.join(tmp, "brawl.inp"), "w").write(txt)
    r = subprocess.run([driver, "dryrun=1"], cwd=tmp, capture_output=True, text=True)
    assert r.returncode != 0 and "Missing 'n_species' in system file" in r.stderr


@pytest.mark.gpu
def test_driver_reproduces_reference_case_01(driver, golden, tmp_path):
    tmp = str(tmp_path)
    write_case(tmp, golden, "01", ("brawl.inp", "metropolis.inp", "fcc_epi.vij"))
    subprocess.run([driver], cwd=tmp, check=True, capture_output=True)
    assert open(os.path.join(tmp, "trajectories/proc_0000_energy_trajectory_at_T_0300.0.dat")).read() == str(golden["t01_energy_txt"])
    assert open(os.path.join(tmp, "trajectories/proc_0000_asro_trajectory_at_T_0300.0.dat")).read() == str(golden["t01_asro_txt"])
    assert np.array_equal(nc_var(os.path.join(tmp, "configs/proc_0000_initial_config_at_T_0300.0.nc"), "configuration")[..., 0], golden["t01_initial"])
    assert np.array_equal(nc_var(os.path.join(tmp, "configs/proc_0000_final_config_at_T_0300.0.nc"), "configuration")[..., 0], golden["t01_final"])
    for k in ("rho", "r", "T", "U"):
        assert np.array_equal(nc_var(os.path.join(tmp, "asro/proc_0000_rho_of_T.nc"), k + " data"), golden["t01_rho_" + k])


@pytest.mark.gpu
def test_driver_reproduces_reference_case_02_four_ranks(driver, golden, tmp_path):
    tmp = str(tmp_path)
    write_case(tmp, golden, "02", ("brawl.inp", "metropolis.inp", "bcc_epi.vij"))
    subprocess.run([driver, "ranks=4"], cwd=tmp, check=True, capture_output=True)
    for r in range(4):
        p = "t02_r%d_" % r
        f = "proc_%04d" % r
        assert open(os.path.join(tmp, "trajectories/%s_energy_trajectory_at_T_0300.0.dat" % f)).read() == str(golden[p + "energy_txt"])
        assert open(os.path.join(tmp, "trajectories/%s_asro_trajectory_at_T_0300.0.dat" % f)).read() == str(golden[p + "asro_txt"])
        assert open(os.path.join(tmp, "energies/%s_energy_diagnostics.dat" % f)).read() == str(golden[p + "diag_txt"])
        assert np.array_equal(nc_var(os.path.join(tmp, "configs/%s_final_config_at_T_0300.0.nc" % f), "configuration")[..., 0], golden[p + "final"])
        for k in ("rho", "r", "T", "U"):
            assert np.array_equal(nc_var(os.path.join(tmp, "asro/%s_rho_of_T.nc" % f), k + " data"), golden[p + "rho_" + k])
    # whole files, byte for byte (NetCDF header incl. attributes + data)
    assert open(os.path.join(tmp, "configs/proc_0000_initial_config_at_T_0300.0.nc"), "rb").read() == golden["raw_t02_r0_initial_nc"].tobytes()
    assert open(os.path.join(tmp, "asro/proc_0000_rho_of_T.nc"), "rb").read() == golden["raw_t02_r0_rho_nc"].tobytes()
    # rank averages (comms_reduce_metropolis_results)
    assert open(os.path.join(tmp, "energies/av_energy_diagnostics.dat")).read() == str(golden["t02_av_diag_txt"])
    assert np.allclose(nc_var(os.path.join(tmp, "asro/av_radial_density.nc"), "rho data"), golden["t02_av_rho_rho"], rtol=0, atol=1e-15)


@pytest.mark.gpu
def test_driver_runs_reference_case_03_nested_sampling(driver, golden, orc, tmp_path):
    """tests/03_serial_nested_sampling through brawl_driver.  The committed golden predates the
    species_numbers quota path (it needs thresholds cum(0.2 l), which no HEAD input
    produces), so the driver -- which follows HEAD (initialise.F90:468-472) -- is checked bit-for-bit
    against the oracle run with HEAD's quotas, and against the golden for the file format.  The golden's
    own numbers are reproduced from the GPU in test_gpu_parity.py::test_nested_sampling_golden_03_on_gpu."""
    tmp = str(tmp_path)
    write_case(tmp, golden, "03", ("brawl.inp", "ns_input.inp", "fcc_al_1.00_crfeconi.vij"))
    subprocess.run([driver], cwd=tmp, check=True, capture_output=True)
    mine = open(os.path.join(tmp, "fcc_al_1.00_crfeconi_K100.energies")).read().split("\n")
    ref = str(golden["t03_energies_txt"]).split("\n")
    assert mine[0] == ref[0] and len(mine) == len(ref)
    for a, b in zip(mine[1:], ref[1:]):      # same field layout (F form with 5 trailing blanks / 3-digit-exponent E form)
        assert len(a) == len(b) and a[:12] == b[:12]
        if a.strip():
            assert ("E" in a) == (abs(float(a.split()[1])) < 0.1)
    sysm = orc.System("fcc", 3, 3, 3, 5, 4, golden["t03_V"])
    conc, cnt = sysm.quotas(numbers=[21, 21, 21, 21, 24])
    culled, _, _ = sysm.nested_sampling(orc.MT(rank=0), conc, cnt, 100, 500, 1000)
    got = np.array([float(l.split()[1]) for l in mine[1:1001]])
    assert np.array_equal(got, culled)


@pytest.mark.gpu
def test_driver_alro_site_occupancies(driver, golden, tmp_path):
    """calculate_alro = T on reference case 01 (metropolis.F90:394-399, 444-447, 506-510; store_state,
    analytics.f90:43-64; ncdf_order_writer, netcdf_io.f90:362-480).  With one ALRO sample at the last trial the file
    must hold the one-hot of the golden final configuration; with a sample every 64 trials the per-site means."""
    tmp = str(tmp_path)
    write_case(tmp, golden, "01", ("brawl.inp", "metropolis.inp", "fcc_epi.vij"))
    inp = str(golden["in_01_metropolis.inp"]).replace("calculate_alro = F", "calculate_alro = T")
    open(os.path.join(tmp, "metropolis.inp"), "w").write(inp)
    subprocess.run([driver], cwd=tmp, check=True, capture_output=True)
    f = netcdf_file(os.path.join(tmp, "alro/proc_0000_rho_of_T.nc"), "r", mmap=False)
    final = golden["t01_final"]
    S = int(getattr(f, "Number of Species"))
    gz, gy, gx = final.shape
    assert S == int(final.max()) and f.N_Basis == 1 and 2 * f.N_1 == gx
    assert list(f.dimensions.items()) == [("b", S), ("x", 1), ("y", gx), ("z", gy), ("s", gz), ("t", 1), ("temp", 1)]
    assert f.variables["grid data"].dimensions == ("t", "s", "z", "y", "x", "b")
    order = np.array(f.variables["grid data"].data)[0, :, :, :, 0, :]          # [z][y][x][species]
    assert np.array_equal(order, np.stack([(final == s) for s in range(1, S + 1)], axis=-1).astype(np.float64))
    assert np.array_equal(np.array(f.variables["temperature data"].data), [300.0])
    # the other outputs are untouched by the extra sampling
    assert np.array_equal(nc_var(os.path.join(tmp, "configs/proc_0000_final_config_at_T_0300.0.nc"), "configuration")[..., 0], final)
    # four samples per temperature, two temperatures: means in [0, 1], one atom per site, accumulator cleared per T
    inp2 = inp.replace("n_sample_steps_alro = 256", "n_sample_steps_alro = 64").replace("T_steps = 1", "T_steps = 2").replace("delta_T = 0.0", "delta_T = 100.0")
    open(os.path.join(tmp, "metropolis.inp"), "w").write(inp2)
    subprocess.run([driver], cwd=tmp, check=True, capture_output=True)
    o2 = nc_var(os.path.join(tmp, "alro/proc_0000_rho_of_T.nc"), "grid data")[:, :, :, :, 0, :]
    assert o2.shape == (2, gz, gy, gx, S) and o2.min() >= 0.0 and o2.max() <= 1.0
    assert np.array_equal(np.unique(o2 * 4), np.unique(np.round(o2 * 4)))       # multiples of 1/4
    for t in range(2):
        assert np.array_equal(o2[t].sum(axis=-1), (final > 0).astype(np.float64))


# ---- Wang-Landau: the C++ host arithmetic of wl_main against the oracle (brawl_host_wl_* hooks, no GPU) -------------------
@pytest.fixture(scope="module")
def hostlib(driver):
    import ctypes as C
    L = C.CDLL(os.path.join(ROOT, "brawl_b200", "host", "libbrawl_host.so"))
    L.brawl_host_wl_replica_exchange.restype = C.c_int
    return L


def _vp(a):
    import ctypes as C
    return a.ctypes.data_as(C.c_void_p)


def test_wl_host_arithmetic_matches_oracle(hostlib, orc):
    """divide_range / create_overlap / create_energy_bins, dos_combine (:1147-1194), mpi_window_optimise (:1211-1311) and
    compute_mean_energy (:457-477) of the C++ wl_main against the oracle's restatements (and the Python mirror for the
    window tables): integers and f64 results identical, exp()-dependent ones to 1e-13."""
    import ctypes as C
    from brawl_b200 import wang_landau as wl
    rng = np.random.default_rng(3)
    for _ in range(200):
        W, bins = int(rng.integers(1, 12)), int(rng.choice([64, 128, 512]))
        ov = float(rng.choice([0.0, 0.1, 0.25, 0.5]))
        iv, idx = np.zeros((W, 2), dtype=np.int64), np.zeros((W, 2), dtype=np.int64)
        hostlib.brawl_host_wl_divide_range(bins, W, C.c_float(ov), _vp(iv), _vp(idx))
        assert np.array_equal(iv, wl.divide_range(bins, W)) and np.array_equal(idx, wl.create_overlap(wl.divide_range(bins, W), ov))
        if W > 1:
            base = np.cumsum(rng.normal(0.3, 1.0, bins))
            lng = np.zeros((W, bins))
            for q in range(W):
                lo, hi = idx[q]
                lng[q, lo - 1:hi] = base[lo - 1:hi] + rng.normal(0, 0.05, hi - lo + 1) + rng.normal(0, 30)
            out = np.zeros(bins)
            hostlib.brawl_host_wl_dos_combine(_vp(lng), _vp(idx), W, bins, _vp(out))
            assert np.array_equal(out, orc.wl_dos_combine(lng, idx))
        if W > 1 and max(int(0.02 * bins), 2) * W <= bins:
            prev = np.full(W, float(1.0 / np.float32(W)), dtype=np.float64)
            iv2 = iv.copy()
            for it in range(3):
                mc = np.floor(rng.uniform(1, 500, W)) * 100 * 128
                b_iv, b_prev = orc.wl_window_optimise(it, iv2, mc, prev, bins)
                a_iv, a_prev = iv2.copy(), prev.copy()
                hostlib.brawl_host_wl_window_optimise(it, W, _vp(a_iv), _vp(mc), _vp(a_prev), bins)
                assert np.array_equal(a_iv, b_iv) and np.array_equal(a_prev, b_prev)
                iv2, prev = a_iv, a_prev
    edges = np.zeros(513)
    hostlib.brawl_host_wl_energy_bins(128, C.c_float(-96.0), C.c_float(0.0), 512, _vp(edges))
    assert np.array_equal(edges, wl.create_energy_bins(128, -96.0, 0.0, 512))
    lng = np.cumsum(rng.uniform(0, 1, 512))
    width = wl.energy_bin_width(128, -96.0, 0.0, 512)
    me = np.zeros((300, 2))
    hostlib.brawl_host_wl_mean_energy(_vp(lng), _vp(edges), 512, C.c_double(width), _vp(me))
    ref = orc.wl_mean_energy(lng, edges, 512, width)
    assert np.array_equal(me[:, 1], ref[:, 1]) and np.allclose(me[:, 0], ref[:, 0], rtol=1e-13, atol=0)


def test_wl_replica_exchange_matches_oracle(hostlib, orc):
    """replica_exchange (:1392-1519) on the walkers' own MT19937 streams: rank 0 shuffles the candidate rows, the lower
    walker of a matched pair draws the acceptance uniform.  C++ host vs oracle: same exchanges in the same order and the
    same final state of every stream, for random energies / ln g tables / window layouts."""
    import ctypes as C
    from brawl_b200 import wang_landau as wl
    rng = np.random.default_rng(17)
    n_ex = 0
    for case in range(60):
        W, walkers, bins = int(rng.integers(2, 7)), int(rng.integers(1, 9)), 128
        idx = wl.create_overlap(wl.divide_range(bins, W), float(rng.choice([0.1, 0.25, 0.5])))
        edges = wl.create_energy_bins(128, -96.0, 0.0, bins)
        P = W * walkers
        e = np.zeros(P)
        for r in range(P):                                   # a walker somewhere inside its own window
            lo, hi = idx[r // walkers]
            b = int(rng.integers(lo, hi + 1))
            e[r] = edges[b - 1] + rng.uniform(0.05, 0.95) * (edges[b] - edges[b - 1])
        lng_w = np.cumsum(rng.normal(0.2, 1.0, (W, bins)), axis=1)
        lng = np.repeat(lng_w, walkers, axis=0)
        MTs = (orc.MT * P)()
        st = np.zeros((P, 625), dtype=np.uint32)
        for r in range(P):
            m = orc.MT(rank=r)
            for _ in range(case):
                m.genrand()
            MTs[r].mt[:] = m.mt[:]; MTs[r].mti = m.mti
            st[r] = m.state625()
        ref = orc.wl_replica_exchange(e, lng, idx, walkers, edges, MTs)
        pairs = np.zeros((P, 2), dtype=np.int32)
        n = hostlib.brawl_host_wl_replica_exchange(_vp(e), _vp(lng), _vp(idx), W, walkers, _vp(edges), bins, _vp(st), _vp(pairs))
        assert [tuple(int(v) for v in p) for p in pairs[:n]] == ref
        for r in range(P):
            assert np.array_equal(st[r, :624], np.frombuffer(MTs[r].mt, dtype=np.uint32)) and int(st[r, 624]) == MTs[r].mti
        for a, b in ref:                                     # exchanges only between adjacent windows
            assert b // walkers == a // walkers + 1
        n_ex += n
    assert n_ex > 20


@pytest.mark.gpu
def test_driver_reproduces_reference_case_04_wang_landau(driver, golden, tmp_path):
    """Mode 302 through the binary: the reference's regression case 04 (4 windows, overlap 0.25, f 0.05 -> 5e-5; ranks=32 =
    mpirun -np 32: eight walkers per window) from its own input files; data/wl_dos.nc against the golden ln g(E) with the
    reference's criterion (NRMSE < 1 %, tests/ci_test.py:42-50), data/wl_dos_bins.nc = create_energy_bins exactly."""
    from brawl_b200 import wang_landau as wl
    tmp = str(tmp_path)
    write_case(tmp, golden, "04", ("brawl.inp", "wl_input.inp", "bcc_epi.vij"))
    r = subprocess.run([driver, "ranks=32"], cwd=tmp, check=True, capture_output=True, text=True)
    assert "Simulation Complete!" in r.stdout and "4 windows x 8 walkers on 1 GPU(s)" in r.stdout
    lng = nc_var(os.path.join(tmp, "data/wl_dos.nc"), "grid data")
    ref = np.asarray(golden["t04_wl_dos"], dtype=np.float64)
    err = float(np.sqrt(np.mean((ref - lng) ** 2)) / np.mean(np.abs(ref)))
    print("brawl_driver mode 302: NRMSE %.4f; %s" % (err, r.stdout.strip().split("\n")[-1]))
    assert lng.shape == (512,) and lng.min() == 0.0 and err < 0.01
    assert np.array_equal(nc_var(os.path.join(tmp, "data/wl_dos_bins.nc"), "grid data"), wl.create_energy_bins(128, -96.0, 0.0, 512))
    assert nc_var(os.path.join(tmp, "data/wl_hist.nc"), "grid data").shape == (512,)
    # mpirun -np 6 with 4 windows: the reference prints the error and exits with status 0 (wang-landau.F90:121-131)
    r = subprocess.run([driver, "ranks=6"], cwd=tmp, capture_output=True, text=True)
    assert r.returncode == 0 and "Number of MPI processes not divisible by num_windows" in r.stdout


@pytest.mark.gpu
def test_driver_wang_landau_two_gpus_through_the_abi_communicator(driver, golden, tmp_path):
    """The same case with the windows sharded over two processes / GPUs (gpus=2): every collective of wl_main goes through the
    C ABI's NCCL communicator (brawl_cuda_comm_*, exchange_replicas).  Needs two devices."""
    import ctypes as C
    n = C.c_int(0)
    lib = C.CDLL(os.path.join(ROOT, "brawl_b200", "libbrawl_cuda.so"))
    if lib.brawl_cuda_device_count(C.byref(n)) != 0 or n.value < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    tmp = str(tmp_path)
    write_case(tmp, golden, "04", ("brawl.inp", "wl_input.inp", "bcc_epi.vij"))
    r = subprocess.run([driver, "ranks=32", "gpus=2"], cwd=tmp, check=True, capture_output=True, text=True, timeout=600)
    assert "4 windows x 8 walkers on 2 GPU(s)" in r.stdout
    lng = nc_var(os.path.join(tmp, "data/wl_dos.nc"), "grid data")
    ref = np.asarray(golden["t04_wl_dos"], dtype=np.float64)
    err = float(np.sqrt(np.mean((ref - lng) ** 2)) / np.mean(np.abs(ref)))
    print("brawl_driver mode 302, 2 GPUs: NRMSE %.4f; %s" % (err, r.stdout.strip().split("\n")[-1]))
    assert err < 0.01


def _feni_inputs(tmp, golden, mode="simulated_annealing", extra=""):
    V = golden["ex_FeNi_V"][: 2 * 2 * 4]
    open(os.path.join(tmp, "FeNi.vij"), "w").write("\n".join(" ".join("%.17g" % v for v in V[i:i + 2]) for i in range(0, V.size, 2)) + "\n")
    open(os.path.join(tmp, "brawl.inp"), "w").write(
        "mode=301\nlattice = fcc\nlattice_parameter = 3.57\nn_1=4\nn_2=4\nn_3=4\nn_species = 2\nspecies_names = Fe Ni\n"
        "species_concentrations = 0.5 0.5\ninteraction_range = 4\ninteraction_file = 'FeNi.vij'\nwc_range = 3\nstatic_seed = .true.\n")
    open(os.path.join(tmp, "metropolis.inp"), "w").write(
        "mode = %s\nn_mc_steps = 25600\nburn_in_start = T\nburn_in = T\nn_burn_in_steps = 2560\nn_sample_steps = 256\n"
        "calculate_energies = T\nn_sample_steps_trajectory = 256\nwrite_trajectory_energy = F\nwrite_trajectory_asro = F\n"
        "calculate_asro = T\nn_sample_steps_asro = 2560\ncalculate_alro = F\nn_sample_steps_alro = 2560\nwrite_trajectory_xyz = F\n"
        "write_final_config_xyz = F\nwrite_final_config_nc = F\nread_start_config_nc = F\nT = 1200\ndelta_T = -100\nT_steps = 13\n"
        "nbr_swap = F\n%s" % (mode, extra))
    return V


@pytest.mark.gpu
def test_driver_feni_annealing_ladder_matches_oracle(driver, orc, golden, tmp_path):
    """BASELINE configs[0] shape (examples/01_metropolis_FeNi/02_simulated_annealing: fcc n = 4, FeNi, 4 shells, burn-in at
    every temperature) on a shortened ladder 1200 K -> 0 K in 13 steps (the last step has beta = +Inf: only dE < 0 is
    accepted, dE = 0 gives NaN and is rejected) through brawl_driver on the reference's MT stream: <E>(T), C(T) and the
    acceptance rate of every temperature equal the oracle's metropolis_simulated_annealing to the printed digits."""
    tmp = str(tmp_path)
    V = _feni_inputs(tmp, golden)
    subprocess.run([driver], cwd=tmp, check=True, capture_output=True)
    rows = [l.split() for l in open(os.path.join(tmp, "energies/proc_0000_energy_diagnostics.dat")).read().strip().split("\n")[1:]]
    assert len(rows) == 13
    sysm = orc.System("fcc", 4, 4, 4, 2, 4, V)
    mt = orc.MT(rank=0)
    conc, cnt = sysm.quotas(conc=[0.5, 0.5])
    g = sysm.initial_setup(mt, conc, cnt)
    for j, row in enumerate(rows):
        temp = 1200.0 - 100.0 * j
        with np.errstate(divide="ignore"):
            beta = 1.0 / (np.float64(temp) * orc.K_B_IN_RY)
        sysm.metropolis_trials(g, mt, beta, 2560)
        e, out = sysm.metropolis_sample(g, mt, temp, 25600, 256)
        assert float(row[0]) == temp
        assert abs(float(row[1]) - out[0]) <= 5e-11 * max(1.0, abs(out[0])) + 1e-10, (j, row, out)
        assert abs(float(row[3]) - out[2]) <= 5.1e-5, (j, row, out)                # printed with four decimals
    assert float(rows[-1][3]) < float(rows[0][3])                       # colder: fewer accepted moves
    assert float(rows[-1][1]) < float(rows[0][1])                       # and lower energy


@pytest.mark.gpu
def test_driver_xyz_outputs_and_decorrelated_samples(driver, golden, tmp_path):
    """xyz_writer (write_xyz.f90:39-116) behind write_final_config_xyz and the decorrelated_samples mode
    (metropolis.F90:572-738): particle count, Lattice line, one `name x y z` per atom in the reference's loop order with
    positions = grid index * a / 2; the species sequence equals the final configuration written as NetCDF."""
    tmp = str(tmp_path)
    _feni_inputs(tmp, golden, extra="")
    txt = open(os.path.join(tmp, "metropolis.inp")).read().replace("write_final_config_xyz = F", "write_final_config_xyz = T") \
        .replace("write_final_config_nc = F", "write_final_config_nc = T").replace("T_steps = 13", "T_steps = 2")
    open(os.path.join(tmp, "metropolis.inp"), "w").write(txt)
    subprocess.run([driver], cwd=tmp, check=True, capture_output=True)
    lines = open(os.path.join(tmp, "configs/proc_0000_final_config_at_T_1100.0.xyz")).read().split("\n")
    assert int(lines[0]) == 256
    assert lines[1].strip().startswith('Lattice="') and lines[1].strip().endswith('"')
    lat = [float(v) for v in lines[1].split('"')[1].split()]
    assert np.allclose(lat, [14.28, 0, 0, 0, 14.28, 0, 0, 0, 14.28])
    cfg = nc_var(os.path.join(tmp, "configs/proc_0000_final_config_at_T_1100.0.nc"), "configuration")[..., 0]    # [z][y][x]
    atoms = [l.split() for l in lines[2:258]]
    want = [(("Fe", "Ni")[cfg[z, y, x] - 1], x, y, z) for x in range(8) for y in range(8) for z in range(8) if cfg[z, y, x]]
    assert len(atoms) == len(want) == 256
    for a, w in zip(atoms, want):
        assert a[0] == w[0] and np.allclose([float(v) for v in a[1:]], [0.5 * 3.57 * c for c in w[1:]], rtol=0, atol=1e-12)
    # decorrelated samples: burn in down the ladder, then n_mc_steps / n_sample_steps dumps at the last temperature
    tmp2 = str(tmp_path / "dec"); os.makedirs(tmp2)
    _feni_inputs(tmp2, golden, mode="decorrelated_samples")
    txt = open(os.path.join(tmp2, "metropolis.inp")).read().replace("T_steps = 13", "T_steps = 3").replace("n_mc_steps = 25600", "n_mc_steps = 5120") \
        .replace("n_sample_steps = 256", "n_sample_steps = 1280").replace("n_sample_steps_asro = 2560", "n_sample_steps_asro = 1280")
    open(os.path.join(tmp2, "metropolis.inp"), "w").write(txt)
    r = subprocess.run([driver], cwd=tmp2, check=True, capture_output=True, text=True)
    files = sorted(os.listdir(os.path.join(tmp2, "configs")))
    assert files == ["proc_0000_config_%04d_at_T_1000.0.xyz" % k for k in range(1, 5)]
    assert r.stdout.count("Burn-in complete at temperature") == 3 and r.stdout.count("Accepted an additional") == 4
    first = open(os.path.join(tmp2, "configs", files[0])).read().split("\n")
    last = open(os.path.join(tmp2, "configs", files[-1])).read().split("\n")
    assert int(first[0]) == 256 and first[2:258] != last[2:258]                 # the chain moved between samples
    assert sorted(l.split()[0] for l in first[2:258]).count("Fe") == 128
