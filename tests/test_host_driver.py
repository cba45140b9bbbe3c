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
