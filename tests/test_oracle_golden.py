"""Pin the CPU oracle against the reference's own golden files.

CPU-only.  These are the reference's regression cases tests/01..03 (tests/ci_test.py demands
identical text lines and np.array_equal on NetCDF variables); fixtures were converted by
tests/golden/make_golden.py.
"""
import os

import numpy as np
import pytest

import fmt


def run_metropolis_case(orc, V, lattice, n, S, shells, rank, n_trials, conc=None, numbers=None, wc_range=3):
    sysm = orc.System(lattice, n, n, n, S, shells, V)
    mt = orc.MT(rank=rank)
    c, cnt = sysm.quotas(conc=conc, numbers=numbers)
    g = sysm.initial_setup(mt, c, cnt)
    g0 = g.copy()
    shells_r = sysm.lattice_shells(g, wc_range)
    e0 = sysm.total_energy(g)
    asro = [sysm.radial_densities(g, wc_range, shells_r)]
    energies, out = sysm.metropolis_sample(g, mt, 300.0, n_trials, 1)
    return sysm, g0, g, np.concatenate([[e0], energies]), out, shells_r, asro


def test_mt19937_known_answer(orc):
    # first outputs of init_genrand(5489) from the published mt19937ar test vector
    mt = orc.MT(seed=5489)
    assert [mt.int32() for _ in range(5)] == [3499211612, 581869302, 3890346734, 3586334585, 545404204]


def test_mt19937_matches_reference_c(orc):
    ref = orc.ref_mt_lib()
    if ref is None:
        pytest.skip("oracle/_ref/libmt19937ar.so not built (reference tree absent)")
    for rank in (0, 1, 3):
        assert ref.f90_init_genrand(0, rank, 0) == 110179 + 11 * rank
        mt = orc.MT(rank=rank)
        for _ in range(2000):
            assert mt.genrand() == ref.genrand()


def test_golden_01_serial_metropolis(orc, golden):
    sysm, g0, g1, E, out, shells, asro = run_metropolis_case(
        orc, golden["t01_V"], "fcc", 4, 5, 6, 0, 256, numbers=[51, 51, 51, 51, 52], wc_range=3)
    assert np.array_equal(g0, golden["t01_initial"])
    assert np.array_equal(g1, golden["t01_final"])
    assert fmt.energy_trajectory(E) == str(golden["t01_energy_txt"])
    assert abs(out[2] - 0.54296875) < 1e-15
    # rho_of_T.nc: r data, U data, rho data (averaged over the single ASRO sample at step 256)
    assert np.array_equal(shells, golden["t01_rho_r"])
    assert out[0] == golden["t01_rho_U"][0]
    rho = sysm.radial_densities(g1, 3, shells)
    assert np.array_equal(rho, golden["t01_rho_rho"][0])
    # first ASRO trajectory line
    lines = str(golden["t01_asro_txt"]).split("\n")
    assert fmt.asro_line(0, asro[0]) == lines[1]


def test_golden_01_asro_trajectory_every_step(orc, golden):
    sysm = orc.System("fcc", 4, 4, 4, 5, 6, golden["t01_V"])
    mt = orc.MT(rank=0)
    c, cnt = sysm.quotas(numbers=[51, 51, 51, 51, 52])
    g = sysm.initial_setup(mt, c, cnt)
    shells = sysm.lattice_shells(g, 3)
    lines = str(golden["t01_asro_txt"]).split("\n")
    beta = 1.0 / (300.0 * orc.K_B_IN_RY)
    # the writer prints the last *sampled* asro (n_sample_steps_asro=256): step 0 value until step 256
    a0 = sysm.radial_densities(g, 3, shells)
    for step in range(1, 257):
        sysm.mc_step(g, mt, beta)
        a = sysm.radial_densities(g, 3, shells) if step % 256 == 0 else a0
        assert fmt.asro_line(step, a) == lines[1 + step]


@pytest.mark.parametrize("rank", [0, 1, 2, 3])
def test_golden_02_parallel_metropolis(orc, golden, rank):
    sysm, g0, g1, E, out, shells, asro = run_metropolis_case(
        orc, golden["t02_V"], "bcc", 4, 4, 6, rank, 128, conc=[0.25] * 4, wc_range=2)
    p = "t02_r%d_" % rank
    assert np.array_equal(g0, golden[p + "initial"])
    assert np.array_equal(g1, golden[p + "final"])
    assert fmt.energy_trajectory(E) == str(golden[p + "energy_txt"])
    assert fmt.diagnostics([300.0], [out[0]], [out[1]], [out[2]]) == str(golden[p + "diag_txt"])
    assert np.array_equal(shells, golden[p + "rho_r"])
    assert np.array_equal(sysm.radial_densities(g1, 2, shells), golden[p + "rho_rho"][0])
    assert out[0] == golden[p + "rho_U"][0]


def test_golden_02_rank_average(orc, golden):
    # comms_reduce_metropolis_results: MPI_Reduce(SUM) to rank 0 then /p (src/comms.F90:122-160)
    outs, rhos = [], []
    for rank in range(4):
        sysm, g0, g1, E, out, shells, _ = run_metropolis_case(
            orc, golden["t02_V"], "bcc", 4, 4, 6, rank, 128, conc=[0.25] * 4, wc_range=2)
        outs.append(out)
        rhos.append(sysm.radial_densities(g1, 2, shells))
    av = np.sum(outs, axis=0) / 4
    txt = fmt.diagnostics([300.0], [av[0]], [av[1]], [av[2]])
    assert txt == str(golden["t02_av_diag_txt"])
    assert np.allclose(np.sum(rhos, axis=0) / 4, golden["t02_av_rho_rho"][0], rtol=0, atol=1e-15)


def test_golden_03_nested_sampling(orc, golden):
    """1000 culled energies at 17 significant digits => bit-level check of total_energy order,
    dE association and E += dE.  Uses thresholds cum(0.2*l) with quotas (21,21,21,21,24): the
    golden was produced by the older quota path."""
    sysm = orc.System("fcc", 3, 3, 3, 5, 4, golden["t03_V"])
    mt = orc.MT(rank=0)
    conc = np.array([0.0, 0.2, 0.2, 0.2, 0.2, 0.2])
    culled, _, _ = sysm.nested_sampling(mt, conc, [21, 21, 21, 21, 24], 100, 500, 1000)
    lines = str(golden["t03_energies_txt"]).strip("\n").split("\n")
    assert lines[0].split() == ["100", "1", "0", "False", "135"]
    ref = np.array([float(l.split()[1]) for l in lines[1:]])
    assert len(ref) == 1000
    assert np.array_equal(culled, ref)          # float(repr17) round-trips => bit-exact


def test_sro_table_counts_equal_cube_scan(orc, golden):
    """The shell-table pair counts (what the GPU SRO kernel computes) reproduce
    radial_densities' cube scan exactly (analytics.f90:293-404)."""
    for key, lat, S, wc in (("t01_final", "fcc", 5, 3), ("t02_r0_final", "bcc", 4, 2), ("t02_r3_final", "bcc", 4, 2)):
        V = golden["t01_V"] if lat == "fcc" else golden["t02_V"]
        sysm = orc.System(lat, 4, 4, 4, S, 6, V)
        g = np.ascontiguousarray(golden[key])
        shells = sysm.lattice_shells(g, wc)
        rho = sysm.radial_densities(g, wc, shells)
        cnt, sc = sysm.radial_counts(g, wc)
        assert np.array_equal(cnt / sc[None, None, :].astype(np.float64), rho)


def test_shell_table_fixture_matches_generated_includes():
    import json, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    js = json.load(open(os.path.join(root, "tests", "golden", "shell_tables.json")))
    for inc in ("oracle/shell_tables.inc", "brawl_b200/csrc/shell_tables.inc"):
        txt = open(os.path.join(root, inc)).read()
        for lat in ("bcc", "fcc", "sc"):
            body = re.search(r"_%s_off\[\d+\]\[3\] = \{(.*?)\n\};" % lat, txt, re.S).group(1)
            trip = [[int(v) for v in m] for m in re.findall(r"\{(-?\d+),(-?\d+),(-?\d+)\}", body)]
            want = [o for s in sorted(js[lat], key=int) for o in js[lat][s]["offsets"]]
            assert trip == want, (inc, lat)
    if os.path.exists("/root/reference/src/bw_hamiltonian.f90"):
        import subprocess, sys
        assert subprocess.call([sys.executable, os.path.join(root, "oracle/tools/extract_shell_tables.py"), "--check"],
                               stdout=subprocess.DEVNULL) == 0
