"""brawl_b200/inputs.py: the reference's own input files (its regression cases, committed as golden text) parsed by the
Python drivers; quotas against the oracle's restatement of initialise.F90:468-506."""
import os

import numpy as np
import pytest


def _write(tmp, golden, case, files):
    for fn in files:
        open(os.path.join(tmp, fn), "w").write(str(golden["in_%s_%s" % (case, fn)]))


def test_control_file_and_exchange_of_the_reference_cases(golden, orc, tmp_path):
    from brawl_b200 import inputs
    want = {"01": ("fcc", 4, 5, "fcc_epi.vij", "t01_V"), "02": ("bcc", 4, 4, "bcc_epi.vij", "t02_V"),
            "03": ("fcc", 3, 5, "fcc_al_1.00_crfeconi.vij", "t03_V"), "04": ("bcc", 4, 4, "bcc_epi.vij", "t04_V")}
    for case, (lat, n, S, vij, vkey) in want.items():
        d = str(tmp_path / case); os.makedirs(d)
        second = {"01": "metropolis.inp", "02": "metropolis.inp", "03": "ns_input.inp", "04": "wl_input.inp"}[case]
        _write(d, golden, case, ("brawl.inp", second, vij))
        p = inputs.read_control_file(os.path.join(d, "brawl.inp"))
        assert (p["lattice"], p["n_1"], p["n_2"], p["n_3"], p["n_species"], p["interaction_file"]) == (lat, n, n, n, S, vij)
        V = inputs.read_exchange(os.path.join(d, vij), S, p["interaction_range"])
        assert np.array_equal(V, golden[vkey][: S * S * p["interaction_range"]])
        conc, count = inputs.species_quotas(p)
        sysm = orc.System(lat, n, n, n, S, p["interaction_range"], V)
        assert inputs.n_atoms(p) == sysm.n_atoms
        if "species_numbers" in p:
            c_ref, n_ref = sysm.quotas(numbers=p["species_numbers"])
        else:
            c_ref, n_ref = sysm.quotas(conc=p["species_concentrations"][1:])
        assert list(n_ref) == count and np.array_equal(c_ref, conc), case
        s = inputs.netcdf_setup(p)
        assert s["n_species"] == S and len(s["species_concentrations"]) == S + 1


def test_quotas_rounding_matches_oracle(orc, golden):
    from brawl_b200 import inputs
    rng = np.random.default_rng(4)
    for _ in range(200):
        S = int(rng.integers(2, 7))
        n = int(rng.integers(2, 6))
        lat = str(rng.choice(["bcc", "fcc", "simple_cubic"]))
        c = rng.dirichlet(np.ones(S))
        p = dict(lattice=lat, n_1=n, n_2=n + 1, n_3=n, n_species=S, species_concentrations=[0.0] + list(c))
        conc, count = inputs.species_quotas(p)
        sysm = orc.System(lat, n, n + 1, n, S, 1, np.zeros(S * S))
        c_ref, n_ref = sysm.quotas(conc=list(c))
        assert list(n_ref) == count and sum(count) == sysm.n_atoms


def test_quotas_refuse_inputs_the_reference_would_hang_on():
    """species_numbers that do not sum to the number of sites (the reference then loops for ever in its redistribution,
    src/initialise.F90:468-506), and all-zero concentrations."""
    from brawl_b200 import inputs, BrawlCudaError
    base = dict(lattice="bcc", n_1=4, n_2=4, n_3=4, n_species=4)
    assert inputs.species_quotas(dict(base, species_numbers=[32, 32, 32, 32]))[1] == [32, 32, 32, 32]
    with pytest.raises(BrawlCudaError, match="species_numbers sum to 127"):
        inputs.species_quotas(dict(base, species_numbers=[32, 32, 32, 31]))
    with pytest.raises(BrawlCudaError, match="all zero"):
        inputs.species_quotas(dict(base, species_concentrations=[0.0] * 5))


def test_error_messages_match_the_reference(golden, tmp_path):
    from brawl_b200 import inputs, BrawlCudaError
    with pytest.raises(BrawlCudaError, match="Could not find input file"):
        inputs.read_control_file(str(tmp_path / "brawl.inp"))
    txt = "\n".join(l for l in str(golden["in_02_brawl.inp"]).split("\n") if not l.startswith("n_species"))
    open(tmp_path / "brawl.inp", "w").write(txt)
    with pytest.raises(BrawlCudaError, match="Missing 'n_species' in system file"):
        inputs.read_control_file(str(tmp_path / "brawl.inp"))
    open(tmp_path / "brawl.inp", "w").write(str(golden["in_02_brawl.inp"]) + "\nspecies_numbers = 32 32 32 32\n")
    with pytest.raises(BrawlCudaError, match="cannot specify both"):
        inputs.read_control_file(str(tmp_path / "brawl.inp"))


def test_wang_landau_from_the_reference_case_04_files(golden, orc, tmp_path, monkeypatch):
    """wl_main's set-up from tests/04_parallel_wang-landau's own files (oracle-backed device: no GPU here)."""
    import test_wl_host_logic as t
    from brawl_b200 import inputs, wang_landau as wl
    t._OracleDevice.orc = orc
    monkeypatch.setattr(wl, "Device", t._OracleDevice)
    d = str(tmp_path)
    _write(d, golden, "04", ("brawl.inp", "wl_input.inp", "bcc_epi.vij"))
    drv, p = inputs.wang_landau_from_files(d, walkers=2)
    assert (drv.p.bins, drv.p.num_windows, drv.p.mc_sweeps, drv.p.performance) == (512, 4, 100, 0)
    assert drv.p.wl_f == float(np.float32(0.05)) and drv.p.tolerance == float(np.float32(5e-5))
    assert drv.counts == [32, 32, 32, 32] and drv.n_atoms == 128 and p["wc_range"] == 3
    assert drv.window_indices.tolist() == [[1, 128], [97, 256], [217, 384], [343, 512]]
    assert np.array_equal(drv.edges, wl.create_energy_bins(128, -96.0, 0.0, 512))


def test_metropolis_file_and_annealing_from_the_reference_case_02_files(golden, tmp_path, monkeypatch):
    import test_replica_annealing as t
    from brawl_b200 import inputs, replica_annealing as ra
    d = str(tmp_path)
    _write(d, golden, "02", ("brawl.inp", "metropolis.inp", "bcc_epi.vij"))
    m = inputs.read_metropolis_file(os.path.join(d, "metropolis.inp"))
    # what brawl_driver's dryrun prints for the same file (tests/test_host_driver.py): asro=128 alro=128 traj=1
    assert (m["mode"], m["n_mc_steps"], m["n_sample_steps"], m["n_sample_steps_asro"], m["n_sample_steps_alro"],
            m["n_sample_steps_trajectory"], m["T"], m["T_steps"]) == ("simulated_annealing", 128, 1, 128, 128, 1, 300.0, 1)
    monkeypatch.setattr(ra, "Device", t._OracleChains)
    drv, p, m2 = inputs.replica_annealing_from_files(d, n_replicas=2)
    per, av = drv.run()
    assert per["energies_of_T"].shape == (2, 1) and av["temperature"].tolist() == [300.0] and drv.counts == [32, 32, 32, 32]
    assert av["rho_of_T"].shape == (1, p["wc_range"], 4, 4) and drv.attempted == 2 * 128
    with pytest.raises(inputs.BrawlCudaError, match="Missing 'T' in Metropolis input file"):
        open(tmp_path / "m2.inp", "w").write("mode = simulated_annealing\nn_mc_steps = 10\nn_sample_steps = 1\n")
        inputs.read_metropolis_file(str(tmp_path / "m2.inp"))
