"""Replica-batched simulated annealing (brawl_b200/replica_annealing.py): the reference's rank-parallel Metropolis
(metropolis.F90:193-447 per chain, comms.F90:122-160 for the final average) with replicas as chains.  CPU: control flow
and the world-size-2 reduction with the oracle standing in for the device; GPU: the real thing on a small lattice."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _OracleChains:
    """Host-facing subset of brawl_b200.Device with the oracle running the chains (tests only)."""

    def __init__(self, lattice, n_1, n_2, n_3, n_species, n_shells, V_ex, device=0, n_replicas=1):
        from oracle import oracle
        self.o = oracle
        self.sys = oracle.System(lattice, n_1, n_2, n_3, n_species, n_shells, V_ex)
        self.n_atoms, self.n_replicas, self.dims, self.lattice = self.sys.n_atoms, n_replicas, (n_1, n_2, n_3), lattice
        self.g = [None] * n_replicas
        self.mt = [oracle.MT(seed=5000 + 100 * device + r) for r in range(n_replicas)]
        self.rng = np.random.default_rng(900 + device)

    def random_config(self, species_count, first_replica=0, n=1, seed=0, offset=0):
        from brawl_b200 import wang_landau as wl
        for r in range(first_replica, first_replica + n):
            self.g[r] = wl.random_configuration(self.lattice, *self.dims, species_count, self.rng)

    def metropolis_run(self, beta, n_trials, seed=0, nbr_swap=False, offset=None):
        acc = np.array([self.sys.metropolis_trials(g, mt, beta, n_trials, nbr_swap) for g, mt in zip(self.g, self.mt)], dtype=np.int64)
        return np.full(self.n_replicas, n_trials, dtype=np.int64), acc, np.zeros(self.n_replicas)

    def total_energy(self, first_replica=0, n=1, exact_order=True):
        return np.array([self.sys.total_energy(g) for g in self.g[first_replica:first_replica + n]])

    def radial_densities_batch(self, wc_range, first_replica=0, n=1):
        return np.stack([self.sys.radial_densities(g, wc_range, self.sys.lattice_shells(g, wc_range))
                         for g in self.g[first_replica:first_replica + n]])


def _run(rank, world, golden, torch_device=None):
    from brawl_b200 import replica_annealing as ra
    drv = ra.ReplicaAnnealing("bcc", 4, 4, 4, 4, 6, golden["t04_V"], [32] * 4, n_replicas=3, T=1500.0, T_steps=3, delta_T=-500.0,
                              n_mc_steps=2560, n_sample_steps=128, n_burn_in_steps=1280, burn_in_start=True, burn_in=True,
                              n_sample_steps_asro=640, wc_range=3, device=rank, rank=rank, world=world, torch_device=torch_device,
                              device_cls=_OracleChains)
    return drv, drv.run()


def test_annealing_control_flow_single_rank(golden, tmp_path):
    drv, (per, av) = _run(0, 1, golden)
    from scipy.io import netcdf_file
    from brawl_b200.replica_annealing import save_av_radial_density
    shells = drv.dev.sys.lattice_shells(drv.dev.g[0], 3)
    save_av_radial_density(str(tmp_path), av, shells, dict(n_1=4, n_2=4, n_3=4, n_species=4, lattice="bcc", interaction_file="bcc_epi.vij",
                                                           species_concentrations=[0.0, 0.25, 0.25, 0.25, 0.25], wc_range=3))
    f = netcdf_file(str(tmp_path / "asro" / "av_radial_density.nc"), "r", mmap=False)
    assert f.variables["rho data"].dimensions == ("T", "r", "j", "i") and np.array_equal(f.variables["rho data"].data, av["rho_of_T"])
    assert np.array_equal(f.variables["T data"].data, av["temperature"]) and np.array_equal(f.variables["U data"].data, av["energies_of_T"])
    assert per["energies_of_T"].shape == (3, 3) and per["rho_of_T"].shape == (3, 3, 3, 4, 4)
    assert np.array_equal(av["temperature"], [1500.0, 1000.0, 500.0]) and av["n_chains"] == 3
    for k in ("energies_of_T", "C_of_T", "acceptance_of_T", "rho_of_T"):
        assert np.allclose(av[k], per[k].mean(axis=0), rtol=1e-14, atol=0)
    assert np.all(np.diff(av["energies_of_T"]) < 0)                       # cooling lowers <E>
    assert np.all((per["acceptance_of_T"] > 0) & (per["acceptance_of_T"] <= 1)) and np.all(np.diff(av["acceptance_of_T"]) < 0)
    assert np.all(per["C_of_T"] >= 0)
    assert drv.attempted == 3 * 3 * (1280 + 2560)
    for l, z in enumerate((1, 8, 6)):                                     # 4 ASRO samples per T, each obeying the sum rule
        assert np.allclose(per["rho_of_T"][:, :, l].sum(axis=-2), z) or np.allclose(per["rho_of_T"][:, :, l].sum(axis=-1), z)
    from brawl_b200.replica_annealing import warren_cowley
    a = warren_cowley(av["rho_of_T"][:, 1:], [0.25] * 4, [8, 6])
    assert a.shape == (3, 2, 4, 4) and np.all(np.abs(a) < 4)


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    golden = np.load(os.path.join(ROOT, "tests", "golden", "brawl_golden.npz"))
    import test_replica_annealing as t
    _, (per, av) = t._run(rank, world, golden)
    q.put((rank, per["energies_of_T"].tolist(), av["energies_of_T"].tolist(), av["rho_of_T"].tolist(), av["n_chains"]))
    dist.destroy_process_group()


def test_annealing_world_size_2_gloo():
    """Replicas sharded over two ranks, no data-path collective; the end-of-run average over all six chains is the same
    on both ranks (comms.F90:122-160)."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=180) for _ in ps])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][4] == res[1][4] == 6
    assert res[0][2] == res[1][2] and res[0][3] == res[1][3]
    allchains = np.array(res[0][1] + res[1][1])
    assert not np.array_equal(np.array(res[0][1]), np.array(res[1][1]))     # the ranks ran different chains
    assert np.allclose(res[0][2], allchains.mean(axis=0), rtol=1e-13, atol=0)


@pytest.mark.gpu
def test_annealing_on_gpu(golden):
    """64 quinary replicas of bcc 8^3 cooled 3000 -> 1000 K on the device: batched chains, batched energies, SRO from
    the radial-counts kernel; bookkeeping identities and reproducibility."""
    from brawl_b200 import replica_annealing as ra
    V = golden["ex_AlCrFeCoNi_V"][: 5 * 5 * 4]
    N = 2 * 8 ** 3
    counts = [205, 205, 205, 205, 204]
    kw = dict(n_replicas=64, T=3000.0, T_steps=3, delta_T=-1000.0, n_mc_steps=40 * N, n_sample_steps=4 * N,
              n_burn_in_steps=20 * N, burn_in_start=True, burn_in=True, n_sample_steps_asro=20 * N, wc_range=3)
    outs = []
    for _ in range(2):
        drv = ra.ReplicaAnnealing("bcc", 8, 8, 8, 5, 4, V, counts, **kw)
        outs.append(drv.run())
        assert drv.attempted >= 64 * 3 * 60 * N
    (per, av), (per2, av2) = outs
    assert np.array_equal(per["energies_of_T"], per2["energies_of_T"]) and np.array_equal(av["rho_of_T"], av2["rho_of_T"])
    assert av["n_chains"] == 64 and np.allclose(av["energies_of_T"], per["energies_of_T"].mean(axis=0), rtol=1e-13)
    assert np.all(np.diff(av["energies_of_T"]) < 0) and np.all(np.diff(av["acceptance_of_T"]) < 0)
    assert np.all((per["acceptance_of_T"] > 0.2) & (per["acceptance_of_T"] <= 1.0)) and np.all(per["C_of_T"] >= 0)
    sem = per["energies_of_T"].std(axis=0, ddof=1) / 8.0
    assert np.all(sem < 0.02 * np.abs(av["energies_of_T"]).max())             # the 64 chains agree with each other
    assert len({tuple(row) for row in per["energies_of_T"].round(12).tolist()}) == 64   # and are independent
    g = drv.dev.get_config(0, 64)
    for gi in g[:4]:
        assert np.array_equal(np.bincount(gi.ravel(), minlength=6)[1:], counts)
    for l, z in enumerate((1, 8, 6)):
        assert np.allclose(per["rho_of_T"][:, :, l].sum(axis=-2), z) or np.allclose(per["rho_of_T"][:, :, l].sum(axis=-1), z)


def test_av_energy_diagnostics_text_matches_golden(golden, tmp_path):
    """The reference's own av_energy_diagnostics.dat (case 02) rewritten from the numbers it holds."""
    from brawl_b200.replica_annealing import save_av_energy_diagnostics
    ref = str(golden["t02_av_diag_txt"])
    T, E, C, a = (float(x) for x in ref.split("\n")[1].split())
    save_av_energy_diagnostics(str(tmp_path), dict(temperature=[T], energies_of_T=[E], C_of_T=[C], acceptance_of_T=[a]))
    assert open(tmp_path / "energies" / "av_energy_diagnostics.dat").read() == ref
