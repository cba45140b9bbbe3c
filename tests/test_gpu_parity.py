"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs, and against the reference's golden files.  Bit-exact for energies / dE / replayed
trajectories / pair counts; statistical for the production sampler (tolerances stated inline).
Run on the B200 box with `pytest -m gpu`."""
import ctypes as C

import numpy as np
import pytest

import fmt

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bw():
    import brawl_b200
    brawl_b200.load()
    return brawl_b200


def random_config(orc, sysm, seed, conc=None):
    mt = orc.MT(seed=seed)
    S = sysm.S
    c, cnt = sysm.quotas(conc=[1.0 / S] * S if conc is None else conc)
    return sysm.initial_setup(mt, c, cnt)


def rand_V(S, shells, seed):
    rng = np.random.default_rng(seed)
    V = rng.normal(scale=2e-3, size=(shells, S, S))
    V = 0.5 * (V + V.transpose(0, 2, 1))          # the shipped tables are symmetric
    return np.ascontiguousarray(V).ravel()


# ---------------------------------------------------------------------------------------------
def test_philox_known_answers(bw):
    """Random123 kat_vectors for philox4x32-10."""
    L = bw.load()
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        c = np.array(ctr, dtype=np.uint32); k = np.array(key, dtype=np.uint32); o = np.zeros(4, dtype=np.uint32)
        bw._lib.check(L.brawl_cuda_philox4x32(0, c.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p),
                                              o.ctypes.data_as(C.c_void_p)))
        assert tuple(int(v) for v in o) == want


def test_config_roundtrip_and_validation(bw, orc, golden):
    sysm = orc.System("fcc", 3, 4, 5, 5, 4, golden["t03_V"])
    dev = bw.Device("fcc", 3, 4, 5, 5, 4, golden["t03_V"], n_replicas=3)
    gs = np.stack([random_config(orc, sysm, s) for s in (1, 2, 3)])
    dev.set_config(gs)
    assert np.array_equal(dev.get_config(0, 3), gs)
    dev.copy_replica(2, 0)
    assert np.array_equal(dev.get_config(0, 1), gs[2])
    bad = gs[0].copy(); bad[bad > 0] = 6
    with pytest.raises(bw.BrawlCudaError, match="species"):
        dev.set_config(bad, 1, 1)
    bad = gs[0].copy(); bad[0, 0, 1] = 1          # (x=1,y=0,z=0) is not an fcc site
    with pytest.raises(bw.BrawlCudaError, match="not a lattice site"):
        dev.set_config(bad, 1, 1)
    with pytest.raises(bw.BrawlCudaError):
        dev.set_config(gs, 2, 3)                  # replica range overflow


def test_config_roundtrip_vectorised_converters(bw, orc):
    """Grids whose x extent is a multiple of 16 use the 16-cells-per-thread converters (brw_pack16/unpack16_kernel):
    round trip, several replicas, the energy of the packed lattice, and the same validation errors -- bcc, fcc and
    simple cubic, extents that differ per axis."""
    rng = np.random.default_rng(4)
    for lattice, dims, S in (("bcc", (8, 4, 3), 4), ("fcc", (16, 3, 2), 5), ("simple_cubic", (16, 3, 2), 3)):
        shells = 2
        V = rng.normal(scale=1e-3, size=shells * S * S)
        sysm = orc.System(lattice, *dims, S, shells, V)
        dev = bw.Device(lattice, *dims, S, shells, V, n_replicas=3)
        gs = np.stack([random_config(orc, sysm, s) for s in (11, 12, 13)])
        assert gs.shape[-1] % 16 == 0
        dev.set_config(gs)
        assert np.array_equal(dev.get_config(0, 3), gs)
        assert dev.total_energy(2, 1)[0] == sysm.total_energy(gs[2])
        bad = gs[0].copy(); bad[bad > 0] = S + 1
        with pytest.raises(bw.BrawlCudaError, match="species"):
            dev.set_config(bad, 1, 1)
        if lattice != "simple_cubic":
            bad = gs[0].copy(); bad[0, 0, 1] = 1      # (x=1,y=0,z=0) is not a bcc / fcc site
            with pytest.raises(bw.BrawlCudaError, match="not a lattice site"):
                dev.set_config(bad, 1, 1)


CASES = [("bcc", n) for n in range(1, 11)] + [("fcc", n) for n in range(1, 7)] + [("simple_cubic", 1), ("simple_cubic", 2)]


def test_copy_replicas_batch(bw, orc):
    """brawl_cuda_copy_replicas_batch (walker cloning of batched nested-sampling runs, nested_sampling.f90:151): all
    pairs in one launch, same result as copy_replica pair by pair; a replica may be copied onto itself; a destination that
    is listed twice or is also a source is refused."""
    sysm = orc.System("fcc", 3, 3, 3, 5, 2, np.zeros(50))
    R = 8
    gs = np.stack([random_config(orc, sysm, 60 + r) for r in range(R)])
    dev = bw.Device("fcc", 3, 3, 3, 5, 2, np.zeros(50), n_replicas=R)
    dev.set_config(gs)
    dev.copy_replicas_batch([0, 3, 5, 6], [1, 2, 5, 7])
    want = gs.copy()
    want[1], want[2], want[7] = gs[0], gs[3], gs[6]
    assert np.array_equal(dev.get_config(0, R), want)
    for src, dst in (([0, 1], [2, 2]), ([0, 2], [2, 3]), ([0], [R])):
        with pytest.raises(bw.BrawlCudaError):
            dev.copy_replicas_batch(src, dst)
    dev.copy_replicas_batch([], [])
    assert np.array_equal(dev.get_config(0, R), want)


def test_compact_lattice_host_buffers(bw, orc):
    """brawl_cuda_set_lattice / get_lattice: the compact host form (1 B per atom, species 0..S-1, device site order) is
    the reference's config grid without the empty cells, in the same z, y, x order -- bcc, fcc and sc, several replicas;
    a species byte >= S is refused."""
    for lattice, dims, S in (("bcc", (6, 4, 8), 4), ("fcc", (4, 6, 5), 5), ("simple_cubic", (5, 4, 3), 3)):
        n1, n2, n3 = dims
        sysm = orc.System(lattice, n1, n2, n3, S, 2, np.zeros(S * S * 2))
        R = 3
        gs = np.stack([random_config(orc, sysm, 30 + r) for r in range(R)])
        dev = bw.Device(lattice, n1, n2, n3, S, 2, np.zeros(S * S * 2), n_replicas=R)
        dev.set_config(gs)
        lat = dev.get_lattice(0, R)
        assert lat.shape == (R, sysm.n_atoms)
        for r in range(R):
            assert np.array_equal(lat[r], gs[r][gs[r] > 0].astype(np.uint8) - 1)
        rolled = np.ascontiguousarray(lat[::-1])
        dev.set_lattice(rolled)
        assert np.array_equal(dev.get_config(0, R), gs[::-1])
        dev.set_lattice(lat[1], first_replica=2, n=1)
        assert np.array_equal(dev.get_config(2, 1), gs[1])
        bad = lat[0].copy()
        bad[sysm.n_atoms // 2] = S
        with pytest.raises(bw.BrawlCudaError):
            dev.set_lattice(bad, first_replica=0, n=1)


@pytest.mark.parametrize("lattice,shells", CASES)
def test_site_and_total_energy_bit_exact(bw, orc, lattice, shells):
    """nbr_energy on every site and total_energy in reference order: bit-exact (incl. the buggy
    bcc shells 8 and 10 and the simple-cubic modulo-n wrap)."""
    S = 4
    n = (6, 5, 7)
    V = rand_V(S, shells, 100 + shells)
    sysm = orc.System(lattice, n[0], n[1], n[2], S, shells, V)
    g = random_config(orc, sysm, 42 + shells)
    dev = bw.Device(lattice, n[0], n[1], n[2], S, shells, V)
    dev.set_config(g)
    assert np.array_equal(dev.site_energies(), sysm.site_energies(g))
    e_ref = sysm.total_energy(g)
    assert dev.total_energy(exact_order=True)[0] == e_ref
    # deterministic tree sum: same value up to f64 rounding of a different association
    assert abs(dev.total_energy(exact_order=False)[0] - e_ref) <= 1e-12 * max(1.0, abs(e_ref))


def test_nbr_energy_single_site_bit_exact(bw, orc):
    """brawl_cuda_nbr_energy == setup%nbr_energy for single sites, every lattice, incl. the wrapped
    corner sites and the centre-species override, bit for bit; off-lattice cells and bad arguments fail."""
    for lattice, shells, S in (("bcc", 6, 4), ("fcc", 4, 5), ("simple_cubic", 2, 3)):
        V = rand_V(S, shells, 9)
        sysm = orc.System(lattice, 4, 5, 6, S, shells, V)
        g = random_config(orc, sysm, 5)
        dev = bw.Device(lattice, 4, 5, 6, S, shells, V)
        dev.set_config(g)
        sites = np.argwhere(g > 0)
        rng = np.random.default_rng(1)
        pick = np.concatenate([sites[:3], sites[-3:], sites[rng.integers(0, len(sites), 30)]])
        for z, y, x in pick:
            assert dev.nbr_energy(x, y, z) == sysm.nbr_energy(g, int(x), int(y), int(z))
        z, y, x = (int(v) for v in sites[7])
        other = 1 + (int(g[z, y, x]) % S)
        g2 = g.copy(); g2[z, y, x] = other
        assert dev.nbr_energy(x, y, z, species=other) == sysm.nbr_energy(g2, x, y, z)
        with pytest.raises(bw.BrawlCudaError):
            dev.nbr_energy(99, 0, 0)
        if lattice != "simple_cubic":
            off = np.argwhere(g == 0)[0]
            with pytest.raises(bw.BrawlCudaError):
                dev.nbr_energy(int(off[2]), int(off[1]), int(off[0]))


def test_total_energy_small_and_tiny_boxes(bw, orc, golden):
    """n=1 and n=2 boxes, where neighbour offsets wrap more than once."""
    for lattice, V, S, shells in (("bcc", golden["t02_V"], 4, 6), ("fcc", golden["t01_V"], 5, 6)):
        for n in (1, 2):
            sysm = orc.System(lattice, n, n, n, S, shells, V)
            c = np.zeros(S + 1); c[1:] = 1.0 / S
            mt = orc.MT(seed=5)
            cc, cnt = sysm.quotas(conc=c[1:])
            g = sysm.initial_setup(mt, cc, cnt)
            dev = bw.Device(lattice, n, n, n, S, shells, V)
            dev.set_config(g)
            assert dev.total_energy()[0] == sysm.total_energy(g)
            assert np.array_equal(dev.site_energies(), sysm.site_energies(g))


def test_total_energy_batched_replicas(bw, orc, golden):
    V = golden["t02_V"]
    sysm = orc.System("bcc", 4, 4, 4, 4, 6, V)
    gs = np.stack([random_config(orc, sysm, 10 + r) for r in range(37)])
    dev = bw.Device("bcc", 4, 4, 4, 4, 6, V, n_replicas=37)
    dev.set_config(gs)
    e = dev.total_energy(0, 37, exact_order=True)
    assert np.array_equal(e, np.array([sysm.total_energy(g) for g in gs]))
    e2 = dev.total_energy(5, 10, exact_order=False)
    assert np.allclose(e2, e[5:15], rtol=0, atol=1e-13)


@pytest.mark.parametrize("lattice,shells,S", [("bcc", 4, 4), ("bcc", 6, 4), ("bcc", 10, 3), ("fcc", 4, 5), ("fcc", 6, 5), ("simple_cubic", 2, 3)])
def test_pair_dE_bit_exact(bw, orc, lattice, shells, S):
    V = rand_V(S, shells, 7)
    n = (4, 5, 3)
    sysm = orc.System(lattice, *n, S, shells, V)
    g = random_config(orc, sysm, 99)
    sites = np.flatnonzero(g.ravel() > 0).astype(np.int32)
    rng = np.random.default_rng(3)
    i1 = rng.choice(sites, 4000)
    i2 = rng.choice(sites, 4000)
    # force first-shell neighbours and identical sites into the batch (mutual-neighbour case)
    offs = sysm.offsets()
    gz, gy, gx = g.shape
    for t in range(200):
        x, y, z = i1[t] % gx, (i1[t] // gx) % gy, i1[t] // (gx * gy)
        o = offs[t % len(offs)]
        i2[t] = (((z + o[2]) % gz) * gy + (y + o[1]) % gy) * gx + (x + o[0]) % gx
    i2[200:220] = i1[200:220]
    dev = bw.Device(lattice, *n, S, shells, V)
    dev.set_config(g)
    got = dev.pair_dE(i1, i2)
    want = sysm.pair_dE(g, i1, i2)
    assert np.array_equal(got, want)
    assert np.array_equal(dev.get_config(), g)           # not modified
    with pytest.raises(bw.BrawlCudaError, match="lattice site"):
        empty = np.flatnonzero(g.ravel() == 0)
        if empty.size == 0:
            raise bw.BrawlCudaError("does not address a lattice site")
        dev.pair_dE([int(empty[0])], [int(sites[0])])


# ---------------------------------------------------------------------------------------------
# deterministic replay against the reference's golden files
def replay_case(bw, orc, V, lattice, n, S, shells, rank, n_trials, conc=None, numbers=None):
    sysm = orc.System(lattice, n, n, n, S, shells, V)
    mt = orc.MT(rank=rank)
    c, cnt = sysm.quotas(conc=conc, numbers=numbers)
    g0 = sysm.initial_setup(mt, c, cnt)
    dev = bw.Device(lattice, n, n, n, S, shells, V)
    dev.set_config(g0)
    e0 = dev.total_energy()[0]
    st = mt.state625()
    beta = 1.0 / (300.0 * bw.K_B_IN_RY)
    acc, e = dev.metropolis_replay(beta, n_trials, st, n_sample_steps=1)
    # oracle continues from the same MT state
    g_ref = g0.copy()
    e_ref, out = sysm.metropolis_sample(g_ref, mt, 300.0, n_trials, 1)
    return dev, sysm, g0, np.concatenate([[e0], e]), acc, st, mt, g_ref, e_ref, out


def test_replay_golden_01(bw, orc, golden):
    dev, sysm, g0, E, acc, st, mt, g_ref, e_ref, out = replay_case(
        bw, orc, golden["t01_V"], "fcc", 4, 5, 6, 0, 256, numbers=[51, 51, 51, 51, 52])
    assert np.array_equal(g0, golden["t01_initial"])
    assert np.array_equal(dev.get_config(), golden["t01_final"])
    assert fmt.energy_trajectory(E) == str(golden["t01_energy_txt"])
    assert np.array_equal(E[1:], e_ref)                       # bit-exact, not just to 1e-10
    assert acc == 139 and acc / 256 == 0.54296875
    assert np.array_equal(st, mt.state625())                  # RNG stream position identical
    rho = dev.radial_densities(3)
    assert np.array_equal(rho, golden["t01_rho_rho"][0])


@pytest.mark.parametrize("rank", [0, 1, 2, 3])
def test_replay_golden_02(bw, orc, golden, rank):
    dev, sysm, g0, E, acc, st, mt, g_ref, e_ref, out = replay_case(
        bw, orc, golden["t02_V"], "bcc", 4, 4, 6, rank, 128, conc=[0.25] * 4)
    p = "t02_r%d_" % rank
    assert np.array_equal(g0, golden[p + "initial"])
    assert np.array_equal(dev.get_config(), golden[p + "final"])
    assert fmt.energy_trajectory(E) == str(golden[p + "energy_txt"])
    assert np.array_equal(E[1:], e_ref)
    assert np.array_equal(st, mt.state625())
    # diagnostics line from GPU energies (host arithmetic of metropolis.F90:412-433)
    step_E = 0.0; step_Esq = 0.0
    for e in E[1:]:
        step_E = step_E + e; step_Esq = step_Esq + e * e
    sim_temp = 300.0 * bw.K_B_IN_RY
    U = step_E / 128 / 128
    Cv = (step_Esq / 128 - (step_E / 128) ** 2) / (sim_temp * 300.0) / 128
    assert fmt.diagnostics([300.0], [U], [max(Cv, 0.0)], [acc / 128.0]) == str(golden[p + "diag_txt"])
    assert np.array_equal(dev.radial_densities(2), golden[p + "rho_rho"][0])


@pytest.mark.parametrize("lattice,shells,nbr", [("bcc", 4, True), ("fcc", 4, True), ("bcc", 10, False), ("simple_cubic", 2, False), ("simple_cubic", 1, True)])
def test_replay_matches_oracle_long(bw, orc, lattice, shells, nbr):
    """5000 trials incl. neighbour-swap mode and odd box shapes: configuration, accept count and
    RNG position identical to the oracle."""
    S = 3
    V = rand_V(S, shells, 21)
    n = (3, 4, 5)
    sysm = orc.System(lattice, *n, S, shells, V)
    g = random_config(orc, sysm, 77)
    dev = bw.Device(lattice, *n, S, shells, V)
    dev.set_config(g)
    mt = orc.MT(seed=2024)
    st = mt.state625()
    beta = 1.0 / (800.0 * bw.K_B_IN_RY)
    acc = dev.metropolis_replay(beta, 5000, st, nbr_swap=nbr)
    acc_ref = sysm.metropolis_trials(g, mt, beta, 5000, nbr_swap=nbr)
    assert acc == acc_ref
    assert np.array_equal(dev.get_config(), g)
    assert np.array_equal(st, mt.state625())
    assert dev.total_energy()[0] == sysm.total_energy(g)


def test_nested_sampling_golden_03_on_gpu(bw, orc, golden):
    """The reference's nested-sampling regression case driven from the host with every walk,
    clone and energy on the GPU: 1000 culled energies bit-identical to the golden file."""
    V = golden["t03_V"]
    K, n_steps, n_iter = 100, 500, 1000
    sysm = orc.System("fcc", 3, 3, 3, 5, 4, V)
    mt = orc.MT(rank=0)
    conc = np.array([0.0, 0.2, 0.2, 0.2, 0.2, 0.2]); cnt = [21, 21, 21, 21, 24]
    dev = bw.Device("fcc", 3, 3, 3, 5, 4, V, n_replicas=K)
    energies = np.zeros(K)
    for w in range(K):                                   # nested_sampling.f90:78-97
        g = sysm.initial_setup(mt, conc, cnt)
        dev.set_config(g, w, 1)
        rnde = mt.genrand()
        energies[w] = dev.total_energy(w, 1)[0] + rnde * float(np.float32(1e-8))
    st = mt.state625()
    n_at = 27 * 5
    extra, n_acc = 0, 0
    culled = np.zeros(n_iter)
    for it in range(1, n_iter + 1):
        i_max = int(np.argmax(energies))
        lim = energies[i_max]
        culled[it - 1] = lim
        if it % int(K / 2.0) == 0 and (n_acc < n_at * 0.05) and extra < n_steps * 100:
            extra += n_steps
        mt.load625(st)
        rnd = mt.genrand()
        st = mt.state625()
        irnd = int(np.ceil(rnd * K))
        dev.copy_replica(irnd - 1, i_max)
        energies[i_max] = energies[irnd - 1]
        energies[i_max], n_acc = dev.ns_walk_replay(energies[i_max], lim, n_steps + extra, st, replica=i_max)
    lines = str(golden["t03_energies_txt"]).strip("\n").split("\n")
    ref = np.array([float(l.split()[1]) for l in lines[1:]])
    assert np.array_equal(culled, ref)


def test_wl_sweeps_replay_matches_oracle(bw, orc, golden):
    V = golden["t04_V"]
    sysm = orc.System("bcc", 4, 4, 4, 4, 6, V)
    g = random_config(orc, sysm, 5)
    bins = 512
    edges = sysm.wl_bin_edges(-96.0, 0.0, bins)
    # a random start lies above energy_max; the reference first walks it into the window
    # (enter_energy_window, wang-landau.F90:643-741) -- here a short oracle anneal does that
    sysm.metropolis_trials(g, orc.MT(seed=1), 1.0 / (1500.0 * orc.K_B_IN_RY), 40 * 128)
    e0 = sysm.total_energy(g)
    assert edges[0] < e0 < edges[-1]
    dev = bw.Device("bcc", 4, 4, 4, 4, 6, V)
    dev.set_config(g)
    mt = orc.MT(seed=99)
    st = mt.state625()
    for (lo, hi) in ((1, 512), (200, 512)):
        lng = np.zeros(bins); hist = np.zeros(hi - lo + 1)
        lng_r = lng.copy(); hist_r = hist.copy()
        acc, ef = dev.wl_sweeps_replay(lng, hist, edges, lo, hi, 0.05000000074505806, 12800, st)
        acc_r, ef_r = sysm.wl_sweeps(g, mt, lng_r, hist_r, edges, lo, hi, 0.05000000074505806, 12800)
        assert acc == acc_r and ef == ef_r
        assert np.array_equal(lng, lng_r) and np.array_equal(hist, hist_r)
        assert np.array_equal(dev.get_config(), g)
        assert np.array_equal(st, mt.state625())


def test_radial_counts_match_oracle(bw, orc, golden):
    for lattice, n, S, wc in (("bcc", (4, 4, 4), 4, 2), ("fcc", (4, 4, 4), 5, 3), ("bcc", (6, 5, 7), 3, 5), ("fcc", (5, 6, 7), 4, 4), ("bcc", (3, 3, 3), 4, 3)):
        V = rand_V(S, 2, 1)
        sysm = orc.System(lattice, *n, S, 2, V)
        g = random_config(orc, sysm, 31)
        dev = bw.Device(lattice, *n, S, 2, V)
        dev.set_config(g)
        shells = sysm.lattice_shells(g, wc)
        rho_ref = sysm.radial_densities(g, wc, shells)
        assert np.array_equal(dev.radial_densities(wc), rho_ref), (lattice, n)


# ---------------------------------------------------------------------------------------------
# production sampler
def test_production_conservation_and_energy_bookkeeping(bw, orc, golden):
    """Concurrent swaps must be non-interacting: the sum of accepted dE equals the change of the
    exact total energy (it would not if two simultaneous trials interacted), and species counts
    are conserved."""
    for lattice, n, S, shells, V, nbr in (("bcc", (16, 16, 16), 4, 4, golden["ex_AlTiCrMo_V"], False),
                                          ("bcc", (16, 12, 20), 4, 6, golden["ex_AlTiCrMo_V"], False),
                                          ("fcc", (12, 12, 12), 5, 4, golden["ex_AlCrFeCoNi_V"], False),
                                          ("fcc", (10, 12, 14), 5, 6, golden["t01_V"], False),
                                          ("bcc", (16, 16, 16), 4, 4, golden["ex_AlTiCrMo_V"], True),
                                          ("fcc", (12, 12, 12), 5, 4, golden["ex_AlCrFeCoNi_V"], True)):
        V = V[: S * S * shells]
        sysm = orc.System(lattice, *n, S, shells, V)
        g = random_config(orc, sysm, 8)
        dev = bw.Device(lattice, *n, S, shells, V)
        plan = dev.metropolis_plan(nbr)
        assert plan["use_box"] >= 1, plan
        dev.set_config(g)
        e0 = dev.total_energy()[0]
        beta = 1.0 / (600.0 * bw.K_B_IN_RY)
        tot_dE = 0.0
        for rep in range(3):
            att, acc, dE = dev.metropolis_run(beta, 20 * sysm.n_atoms, nbr_swap=nbr)
            assert att[0] >= 20 * sysm.n_atoms and 0 < acc[0] <= att[0]
            tot_dE += dE[0]
        g1 = dev.get_config()
        assert np.array_equal(np.bincount(g1.ravel(), minlength=S + 1), np.bincount(g.ravel(), minlength=S + 1))
        assert np.array_equal(g1 == 0, g == 0)
        e1 = sysm.total_energy(g1)
        assert e1 == dev.total_energy()[0]
        assert e1 < e0                                        # 600 K from a random start: energy drops
        assert abs((e1 - e0) - tot_dE) < 1e-10 * abs(e0) + 1e-12, (lattice, n, e0, e1, tot_dE)


def test_production_planner_variants(bw, orc, golden):
    """Three valid decompositions of the same lattice: the dense non-interacting-set plan of the word-lattice
    kernel (default: box 64x64x32, period 4, 896 trials per step), rectangular multi-orientation periods on a
    user box, and cubic periods (test hook).  Energy bookkeeping holds for each (it would not if two simultaneous
    trials interacted); rectangular periods expose more simultaneous trials than cubic ones."""
    V = golden["ex_AlTiCrMo_V"][:64]
    sysm = orc.System("bcc", 32, 32, 32, 4, 4, V)
    g = random_config(orc, sysm, 8)
    plans = []
    for box, steps in (((0, 0, 0), 0), ((32, 32, 32), 0), ((32, 32, 32), 100000)):
        dev = bw.Device("bcc", 32, 32, 32, 4, 4, V)
        dev.metropolis_tune(box, steps)
        plan = dev.metropolis_plan()
        plans.append(plan)
        dev.set_config(g)
        e0 = dev.total_energy()[0]
        att, acc, dE = dev.metropolis_run(1.0 / (700.0 * bw.K_B_IN_RY), 6 * sysm.n_atoms)
        e1 = dev.total_energy()[0]
        # the word kernels sum fixed-point dE (epoch kernel: 22-bit table entries, < 2e-10 Ry per accepted swap)
        tol = 2e-10 * float(acc[0]) if plan["use_box"] == 4 else 1e-10 * abs(e0) + 1e-12
        assert abs((e1 - e0) - dE[0]) < tol
        g1 = dev.get_config()
        assert dev.total_energy()[0] == sysm.total_energy(g1)
        assert np.array_equal(np.bincount(g1.ravel(), minlength=5), np.bincount(g.ravel(), minlength=5))
    assert plans[0]["use_box"] == 4 and plans[0]["P"] == (4, 4, 4) and plans[0]["n_orientations"] == 1
    # default: epoch kernel, 68x64x32 boxes at a pitch of 64x64x28 (shared margin planes in x and z), 32 warps x 30 lanes
    assert (plans[0]["box_x"], plans[0]["box_y"], plans[0]["box_z"], plans[0]["trials_per_step"]) == (64, 64, 28, 960)
    assert plans[2]["P"] == (6, 6, 6) and plans[2]["n_orientations"] == 1
    assert plans[1]["P"][0] * plans[1]["P"][1] * plans[1]["P"][2] < 216 and plans[1]["n_orientations"] == 3
    assert plans[1]["trials_per_step"] > plans[2]["trials_per_step"]


def test_dense_decomposition_conservation_exact(bw, orc, golden):
    """Word-lattice kernel, EXACT instantiation (dE_mode 0: reference association for every trial) on the dense
    decomposition, 4 and 5 species: the sum of accepted dE equals the change of the oracle's total energy to
    rounding -- any two interacting simultaneous trials would break it -- and species counts are conserved."""
    for S, key in ((4, "ex_AlTiCrMo_V"), (5, "ex_AlCrFeCoNi_V")):
        V = golden[key][: S * S * 4]
        sysm = orc.System("bcc", 32, 32, 32, S, 4, V)
        g = random_config(orc, sysm, 21)
        dev = bw.Device("bcc", 32, 32, 32, S, 4, V)
        dev.metropolis_set_mode(0)
        assert dev.metropolis_plan()["use_box"] == 5
        dev.set_config(g)
        e0 = sysm.total_energy(g)
        tot = 0.0
        for rep in range(2):
            att, acc, dE = dev.metropolis_run(1.0 / (900.0 * bw.K_B_IN_RY), 10 * sysm.n_atoms, seed=5 + rep)
            assert att[0] >= 10 * sysm.n_atoms and 0 < acc[0] < att[0]
            tot += dE[0]
        g1 = dev.get_config()
        assert np.array_equal(np.bincount(g1.ravel(), minlength=S + 1), np.bincount(g.ravel(), minlength=S + 1))
        assert np.array_equal(g1 == 0, g == 0)
        e1 = sysm.total_energy(g1)
        assert e1 < e0
        assert abs((e1 - e0) - tot) < 1e-10 * abs(e1 - e0) + 1e-11, (S, e0, e1, tot)


def test_dense_decomposition_statistics(bw, orc, golden):
    """Equilibrium energy per atom and first/second-shell pair counts of the dense-set sampler (word kernel) agree
    with the period-P sublattice sampler (byte kernels, itself checked against the oracle's sequential sampler in
    test_production_statistics_match_oracle).  T = 2500 K: both samplers are equilibrated after ~50 sweeps (at
    1200 K the alloy is still ordering after 600 sweeps, and the dense sampler -- which exchanges species between
    ALL pairs of residue classes -- orders faster than the period-P sampler, so the comparison there is not an
    equilibrium one).  Tolerance: 5 standard errors from block averages + 2e-6 Ry."""
    V = golden["ex_AlTiCrMo_V"][:64]
    n, T = 32, 2500.0
    sysm = orc.System("bcc", n, n, n, 4, 4, V)
    g = random_config(orc, sysm, 3)
    N = sysm.n_atoms
    stats = []
    for byte_layout in (False, True):
        dev = bw.Device("bcc", n, n, n, 4, 4, V)
        dev.metropolis_set_layout(byte_layout)
        assert dev.metropolis_plan()["use_box"] == (3 if byte_layout else 4)
        dev.set_config(g)
        beta = 1.0 / (T * bw.K_B_IN_RY)
        dev.metropolis_run(beta, 150 * N, seed=11)                # equilibrate
        es, cs = [], []
        for k in range(60):
            dev.metropolis_run(beta, 4 * N, seed=100 + k)
            es.append(dev.total_energy()[0] / N)
            cnt, spc = dev.radial_counts(3)
            cs.append(cnt[1:3].astype(np.float64) / N)
        es, cs = np.array(es), np.array(cs)
        blocks = es.reshape(6, 10).mean(axis=1)
        stats.append((es.mean(), blocks.std(ddof=1) / np.sqrt(6), cs.mean(axis=0), cs.reshape(6, 10, *cs.shape[1:]).mean(axis=1).std(axis=0, ddof=1) / np.sqrt(6)))
    (e0, s0, c0, sc0), (e1, s1, c1, sc1) = stats
    assert abs(e0 - e1) < 5 * np.hypot(s0, s1) + 2e-6, (e0, e1, s0, s1)
    assert np.all(np.abs(c0 - c1) < 5 * np.hypot(sc0, sc1) + 2e-3), (c0, c1)


def test_production_is_deterministic(bw, orc, golden):
    V = golden["ex_AlTiCrMo_V"][:64]
    sysm = orc.System("bcc", 16, 16, 16, 4, 4, V)
    g = random_config(orc, sysm, 8)
    outs = []
    for _ in range(2):
        dev = bw.Device("bcc", 16, 16, 16, 4, 4, V)
        dev.set_config(g)
        res = dev.metropolis_run(1.0 / (900.0 * bw.K_B_IN_RY), 100000, seed=1234)
        outs.append((dev.get_config().copy(), res))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert all(np.array_equal(a, b) for a, b in zip(outs[0][1], outs[1][1]))
    dev = bw.Device("bcc", 16, 16, 16, 4, 4, V)
    dev.set_config(g)
    dev.metropolis_run(1.0 / (900.0 * bw.K_B_IN_RY), 100000, seed=1235)
    assert not np.array_equal(dev.get_config(), outs[0][0])


@pytest.mark.parametrize("lattice,n,S,shells,key", [("bcc", 32, 4, 4, "ex_AlTiCrMo_V"), ("bcc", 32, 4, 6, "t02_V"),
                                                    ("fcc", 32, 5, 4, "ex_AlCrFeCoNi_V"), ("fcc", 32, 5, 6, "t01_V")])
def test_specialised_kernel_equals_generic_kernel(bw, orc, golden, lattice, n, S, shells, key):
    """The compile-time-geometry kernels and the generic runtime-geometry kernel implement the same
    algorithm with the same Philox counters: identical trajectories, bit for bit."""
    V = golden[key][: S * S * shells]
    g = np.zeros((2 * n, 2 * n, 2 * n), dtype=np.int8)
    rng = np.random.default_rng(4)
    z, y, x = np.meshgrid(np.arange(2 * n), np.arange(2 * n), np.arange(2 * n), indexing="ij")
    mask = ((x & 1) == (z & 1)) & ((y & 1) == (z & 1)) if lattice == "bcc" else ((x + y + z) & 1) == 0
    g[mask] = rng.integers(1, S + 1, size=int(mask.sum()))
    res = []
    for generic in (False, True):
        dev = bw.Device(lattice, n, n, n, S, shells, V)
        dev.metropolis_set_layout(True)                 # byte-lattice kernels (the word kernel has its own decomposition)
        if generic:
            dev.metropolis_tune((0, 0, 0), -1)          # automatic steps, generic kernel forced
        plan = dev.metropolis_plan()
        assert plan["use_box"] == (1 if generic else 3), plan
        dev.set_config(g)
        out = dev.metropolis_run(1.0 / (700.0 * bw.K_B_IN_RY), 3 * int(mask.sum()), seed=77)
        res.append((dev.get_config().copy(), out))
    assert np.array_equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][1][0], res[1][1][0]) and np.array_equal(res[0][1][1], res[1][1][1])
    # the specialised kernel screens with count-based dE (same decisions, last-bit different dE sum)
    assert np.allclose(res[0][1][2], res[1][1][2], rtol=0, atol=1e-9)


@pytest.mark.parametrize("lattice,n,S,shells,key,T", [("bcc", 32, 4, 4, "ex_AlTiCrMo_V", 300.0), ("bcc", 32, 4, 4, "ex_AlTiCrMo_V", 2000.0),
                                                      ("bcc", 32, 5, 4, "ex_AlCrFeCoNi_V", 800.0), ("fcc", 32, 5, 4, "ex_AlCrFeCoNi_V", 600.0),
                                                      ("fcc", 32, 2, 6, "t01_V", 500.0), ("bcc", 32, 4, 6, "t02_V", 700.0),
                                                      ("fcc", 32, 2, 4, "ex_FeNi_V", 400.0)])
def test_screened_kernel_trajectory_identical(bw, orc, golden, lattice, n, S, shells, key, T):
    """dE_mode 1 (integer-count screening on the byte lattice) and dE_mode 2 (word lattice, fixed-point dp4a dE,
    ex2.approx acceptance test) recompute every trial inside their guard bands with the reference association,
    so they take exactly the accept/reject decisions of dE_mode 0 (reference association for every trial):
    same seed => identical configuration and identical accept counts after millions of trials."""
    S0 = {"ex_AlTiCrMo_V": 4, "t02_V": 4, "ex_FeNi_V": 2}.get(key, 5)          # species of the stored table
    V = golden[key][: S0 * S0 * shells].reshape(shells, S0, S0)[:, :S, :S].copy().ravel()
    g = np.zeros((2 * n, 2 * n, 2 * n), dtype=np.int8)
    rng = np.random.default_rng(4)
    par = np.arange(2 * n) & 1
    mask = ((par[None, None, :] == par[:, None, None]) & (par[None, :, None] == par[:, None, None])) if lattice == "bcc" else \
        (((par[None, None, :] + par[None, :, None] + par[:, None, None]) & 1) == 0)
    g[mask] = rng.integers(1, S + 1, size=int(mask.sum()))
    res = []
    word = lattice == "bcc" and shells == 4
    # (layout, mode, expected kernel kind); runs with the same layout share one decomposition
    # (default layout on the other geometries: the byte-lattice epoch kernels, 7 = reference association, 6 = cached energies)
    runs = [(True, 0, 2), (True, 1, 3)] + ([(False, 0, 5), (False, 2, 4)] if word else [(False, 0, 7), (False, 2, 6)])
    res = {}
    for byte_layout, mode, kind in runs:
        dev = bw.Device(lattice, n, n, n, S, shells, V)
        dev.metropolis_set_layout(byte_layout)
        dev.metropolis_set_mode(mode)
        assert dev.metropolis_plan()["use_box"] == kind
        dev.set_config(g)
        out = dev.metropolis_run(1.0 / (T * bw.K_B_IN_RY), 12 * int(mask.sum()), seed=99)
        res[(byte_layout, mode)] = (dev.get_config().copy(), out, dev.total_energy()[0])
    pairs = [((True, 0), (True, 1), 1e-9), ((False, 0), (False, 2), 1e-4)]
    for ka, kb, tol in pairs:
        a, b = res[ka], res[kb]
        assert np.array_equal(a[0], b[0])
        assert np.array_equal(a[1][0], b[1][0]) and np.array_equal(a[1][1], b[1][1])
        assert a[2] == b[2]
        # sum of accepted dE: same to rounding (epoch kernel: fixed-point dE from a 22-bit table, < 2e-10 Ry per accepted swap)
        assert abs(a[1][2][0] - b[1][2][0]) < tol


def test_production_limits(bw, orc, golden):
    """beta = 0: everything accepted.  beta -> infinity: only dE < 0 accepted, energy monotone."""
    V = golden["ex_AlTiCrMo_V"][:64]
    sysm = orc.System("bcc", 16, 16, 16, 4, 4, V)
    g = random_config(orc, sysm, 11)
    dev = bw.Device("bcc", 16, 16, 16, 4, 4, V)
    dev.set_config(g)
    att, acc, dE = dev.metropolis_run(0.0, 50000)
    assert att[0] == acc[0]
    e_prev = dev.total_energy()[0]
    for _ in range(3):
        att, acc, dE = dev.metropolis_run(1e300, 50000)
        e = dev.total_energy()[0]
        assert e <= e_prev and dE[0] <= 0.0
        e_prev = e


@pytest.mark.parametrize("lattice,n,S,shells,key,T", [("bcc", 12, 4, 4, "ex_AlTiCrMo_V", 1500.0), ("fcc", 8, 2, 4, "ex_FeNi_V", 900.0)])
def test_production_statistics_match_oracle(bw, orc, golden, lattice, n, S, shells, key, T):
    """Equilibrium <E>/atom and first-shell pair densities from the sublattice-parallel sampler
    agree with the oracle's sequential reference sampler.  Tolerance: 5 standard errors of the
    difference (both errors estimated from blocked samples) + 1e-6 Ry/atom floor."""
    V = golden[key][: S * S * shells]
    sysm = orc.System(lattice, n, n, n, S, shells, V)
    N = sysm.n_atoms
    beta = 1.0 / (T * bw.K_B_IN_RY)
    g = random_config(orc, sysm, 3)
    # oracle chain
    mt = orc.MT(seed=555)
    g_o = g.copy()
    sysm.metropolis_trials(g_o, mt, beta, 60 * N)
    eo, ro = [], []
    shells_r = sysm.lattice_shells(g_o, 2)
    for _ in range(40):
        sysm.metropolis_trials(g_o, mt, beta, 10 * N)
        eo.append(sysm.total_energy(g_o) / N)
        ro.append(sysm.radial_densities(g_o, 2, shells_r)[1])
    # GPU chains: 8 replicas
    R = 8
    dev = bw.Device(lattice, n, n, n, S, shells, V, n_replicas=R)
    assert dev.metropolis_plan()["use_box"] >= 1
    dev.set_config(np.stack([g] * R))
    dev.metropolis_run(beta, 60 * N)
    eg, rg = [], []
    for _ in range(20):
        dev.metropolis_run(beta, 10 * N)
        eg.append(dev.total_energy(0, R, exact_order=False) / N)
        rg.append(np.mean([dev.radial_densities(2, r)[1] for r in range(R)], axis=0))
    eo = np.array(eo); eg = np.array(eg)
    se_o = eo.std(ddof=1) / np.sqrt(len(eo) / 4.0)            # allow for autocorrelation (x4)
    se_g = eg.mean(axis=1).std(ddof=1) / np.sqrt(len(eg) / 4.0)
    diff = abs(eo.mean() - eg.mean())
    assert diff < 5.0 * np.hypot(se_o, se_g) + 1e-6, (eo.mean(), eg.mean(), se_o, se_g)
    ro = np.array(ro); rg = np.array(rg)
    se_r = np.hypot(ro.std(axis=0, ddof=1) / np.sqrt(len(ro) / 4.0), rg.std(axis=0, ddof=1) / np.sqrt(len(rg) / 4.0))
    assert np.all(np.abs(ro.mean(axis=0) - rg.mean(axis=0)) < 5.0 * se_r + 0.02), (ro.mean(axis=0), rg.mean(axis=0))


def test_chain_kernel_small_lattices(bw, orc, golden):
    """Lattices too small for boxes run one sequential chain per replica (reference proposal
    distribution, Philox stream): bookkeeping + statistics vs oracle on the FeNi example shape."""
    V = golden["ex_FeNi_V"][: 2 * 2 * 4]
    sysm = orc.System("fcc", 4, 4, 4, 2, 4, V)
    N = sysm.n_atoms
    R = 64
    g = random_config(orc, sysm, 3)
    dev = bw.Device("fcc", 4, 4, 4, 2, 4, V, n_replicas=R)
    assert dev.metropolis_plan()["use_box"] == 0
    dev.set_config(np.stack([g] * R))
    beta = 1.0 / (800.0 * bw.K_B_IN_RY)
    e0 = dev.total_energy(0, R)
    att, acc, dE = dev.metropolis_run(beta, 200 * N)
    assert np.all(att == 200 * N)
    e1 = dev.total_energy(0, R)
    assert np.allclose(e1 - e0, dE, rtol=0, atol=1e-11)
    eg = []
    for _ in range(10):
        dev.metropolis_run(beta, 20 * N)
        eg.append(dev.total_energy(0, R, exact_order=False).mean() / N)
    mt = orc.MT(seed=9)
    g_o = g.copy()
    sysm.metropolis_trials(g_o, mt, beta, 200 * N)
    eo = []
    for _ in range(200):
        sysm.metropolis_trials(g_o, mt, beta, 20 * N)
        eo.append(sysm.total_energy(g_o) / N)
    eo = np.array(eo); eg = np.array(eg)
    se = np.hypot(eo.std(ddof=1) / np.sqrt(len(eo) / 4.0), eg.std(ddof=1) / np.sqrt(len(eg)))
    assert abs(eo.mean() - eg.mean()) < 5 * se + 1e-6, (eo.mean(), eg.mean(), se)


def test_wl_and_ns_production_kernels(bw, orc, golden):
    """Philox-driven WL sweeps / NS walks for walker batches: invariants of the reference update
    rule (ln g grows by wl_f per trial; hist counts every INT(0.02 N)-th trial; energies stay inside
    the window / below the NS ceiling; running energy equals the exact total energy to rounding)."""
    V = golden["t04_V"]
    sysm = orc.System("bcc", 4, 4, 4, 4, 6, V)
    W, bins = 48, 512
    edges = sysm.wl_bin_edges(-96.0, 0.0, bins)
    gs = np.stack([random_config(orc, sysm, 100 + w) for w in range(W)])
    dev = bw.Device("bcc", 4, 4, 4, 4, 6, V, n_replicas=W)
    dev.set_config(gs)
    lng0 = np.zeros((W, bins)); hist0 = np.zeros((W, bins))
    with pytest.raises(bw.BrawlCudaError, match="outside"):      # random starts lie above energy_max
        dev.wl_sweeps(lng0, hist0, edges, 1, bins, 0.05, 10, seed=5)
    for g in gs:
        sysm.metropolis_trials(g, orc.MT(seed=1), 1.0 / (1500.0 * orc.K_B_IN_RY), 40 * 128)
    dev.set_config(gs)
    lng = np.zeros((W, bins)); hist = np.zeros((W, bins))
    n_trials = 12800
    acc, ef = dev.wl_sweeps(lng, hist, edges, 1, bins, 0.05, n_trials, seed=5)
    assert np.allclose(lng.sum(axis=1), 0.05 * n_trials, rtol=1e-12)
    assert np.array_equal(hist.sum(axis=1), np.full(W, n_trials // 2))
    e_exact = dev.total_energy(0, W)
    assert np.allclose(ef, e_exact, rtol=0, atol=1e-11)
    assert np.all((ef >= edges[0]) & (ef < edges[-1]))
    assert np.all(acc > 0)
    # NS: walk under a ceiling
    lim = e_exact + 1e-4
    e2, nacc = dev.ns_walk(np.arange(W), e_exact, lim, 500, seed=6)
    assert np.all(e2 < lim) and np.all(nacc > 0)
    assert np.allclose(e2, dev.total_energy(0, W), rtol=0, atol=1e-11)


def test_walker_kernels_fast_dE_takes_the_reference_decisions(bw, orc, golden):
    """The walker kernels' default lane-parallel dE (butterfly sum) against their reference-association instantiation
    (dE_mode 0) on the same Philox streams: WL sweeps (bcc n=4, 6 shells), NS walks and small-lattice Metropolis chains
    (fcc n=4) must take the same decisions -- identical ln g, histograms, accept counts and final configurations --
    and their running energies agree to f64 rounding."""
    V = golden["t04_V"]
    sysm = orc.System("bcc", 4, 4, 4, 4, 6, V)
    W, bins = 24, 512
    edges = sysm.wl_bin_edges(-96.0, 0.0, bins)
    gs = np.stack([random_config(orc, sysm, 300 + w) for w in range(W)])
    for g in gs:
        sysm.metropolis_trials(g, orc.MT(seed=2), 1.0 / (1500.0 * orc.K_B_IN_RY), 40 * 128)
    res = []
    for mode in (2, 0):
        dev = bw.Device("bcc", 4, 4, 4, 4, 6, V, n_replicas=W)
        dev.metropolis_set_mode(mode)
        dev.set_config(gs)
        lng = np.zeros((W, bins)); hist = np.zeros((W, bins))
        acc, ef = dev.wl_sweeps(lng, hist, edges, 1, bins, 0.05, 12800, seed=9)
        e_exact = dev.total_energy(0, W)
        lim = e_exact + 2e-4
        e2, nacc = dev.ns_walk(np.arange(W), e_exact, lim, 400, seed=10)
        res.append((lng, hist, acc, ef, e2, nacc, dev.get_config(0, W)))
    a, b = res
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert np.allclose(a[3], b[3], rtol=0, atol=1e-11)
    assert np.array_equal(a[5], b[5]) and np.allclose(a[4], b[4], rtol=0, atol=1e-11)
    assert np.array_equal(a[6], b[6])
    # Metropolis chains on a lattice too small for boxes
    Vf = golden["ex_FeNi_V"][: 2 * 2 * 4]
    sf = orc.System("fcc", 4, 4, 4, 2, 4, Vf)
    g = random_config(orc, sf, 5)
    out = []
    for mode in (2, 0):
        dev = bw.Device("fcc", 4, 4, 4, 2, 4, Vf, n_replicas=16)
        dev.metropolis_set_mode(mode)
        assert dev.metropolis_plan()["use_box"] == 0
        dev.set_config(np.stack([g] * 16))
        att, acc, dE = dev.metropolis_run(1.0 / (900.0 * bw.K_B_IN_RY), 100 * sf.n_atoms, seed=3)
        out.append((acc, dE, dev.get_config(0, 16)))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][2], out[1][2])
    assert np.allclose(out[0][1], out[1][1], rtol=0, atol=1e-11)


# ---------------------------------------------------------------------------------------------
# edge cases and full-size properties
def test_edge_cases_empty_and_degenerate_inputs(bw, orc, golden):
    V = golden["t02_V"]
    sysm = orc.System("bcc", 4, 4, 4, 4, 6, V)
    g = random_config(orc, sysm, 2)
    dev = bw.Device("bcc", 4, 4, 4, 4, 6, V)
    dev.set_config(g)
    assert dev.pair_dE([], []).size == 0                                   # empty batch
    st = orc.MT(seed=1).state625(); st0 = st.copy()
    assert dev.metropolis_replay(1.0, 0, st) == 0 and np.array_equal(st, st0)     # zero trials: RNG untouched
    att, acc, dE = dev.metropolis_run(1.0, 0)
    assert att[0] == 0 and acc[0] == 0 and dE[0] == 0.0
    assert np.array_equal(dev.get_config(), g)
    bad = st.copy(); bad[624] = 700
    with pytest.raises(bw.BrawlCudaError, match="MT19937"):
        dev.metropolis_replay(1.0, 10, bad)
    with pytest.raises(bw.BrawlCudaError):
        dev.total_energy(0, 2)                                              # more replicas than the handle owns
    # single-species lattice: every proposal is a same-species "accept", energy is a constant
    V1 = np.array([1e-3, -2e-3, 5e-4, 1e-4])
    one = np.where(g > 0, 1, 0).astype(np.int8)
    d1 = bw.Device("bcc", 4, 4, 4, 1, 4, V1)
    d1.set_config(one)
    assert d1.total_energy()[0] == orc.System("bcc", 4, 4, 4, 1, 4, V1).total_energy(one)
    att, acc, dE = d1.metropolis_run(5.0, 1000)
    assert att[0] == acc[0] == 1000 and dE[0] == 0.0
    # maximum species count of the library (16)
    S = 16
    V16 = rand_V(S, 2, 3)
    s16 = orc.System("fcc", 3, 3, 3, S, 2, V16)
    g16 = random_config(orc, s16, 9)
    d16 = bw.Device("fcc", 3, 3, 3, S, 2, V16)
    d16.set_config(g16)
    assert d16.total_energy()[0] == s16.total_energy(g16)
    with pytest.raises(bw.BrawlCudaError, match="n_species"):
        bw.Device("fcc", 3, 3, 3, 17, 2, np.zeros(17 * 17 * 2))


def test_tiled_total_energy_matches_reference_order(bw, orc):
    """Production total_energy (exact_order=False) on lattices that are a whole number of 32x16x8 tiles runs
    brw_energy_tile_kernel (integer neighbour counts x V); it must agree with the bit-exact reference-order sum, and
    with the oracle's total_energy, to f64 rounding -- bcc and fcc, 4 and 6 shells, 2..5 species, unsymmetric V so a
    centre/neighbour index mix-up would show."""
    rng = np.random.default_rng(17)
    for lattice, n, S, shells in (("bcc", 32, 4, 4), ("bcc", 32, 5, 6), ("bcc", 32, 2, 4), ("fcc", 32, 5, 4), ("fcc", 32, 3, 6)):
        V = rng.normal(scale=2e-3, size=shells * S * S)
        sysm = orc.System(lattice, n, n, n, S, shells, V)
        g = random_config(orc, sysm, 31 + S)
        dev = bw.Device(lattice, n, n, n, S, shells, V)
        dev.set_config(g)
        e_exact = dev.total_energy(exact_order=True)[0]
        e_tile = dev.total_energy(exact_order=False)[0]
        assert e_exact == sysm.total_energy(g)
        assert abs(e_tile - e_exact) <= 1e-12 * abs(e_exact) + 1e-13, (lattice, S, shells, e_tile, e_exact)
        dev.metropolis_tune((0, 0, 0), -1)                       # test hook: generic kernels (row energy kernel)
        assert abs(dev.total_energy(exact_order=False)[0] - e_tile) <= 1e-12 * abs(e_exact) + 1e-13


@pytest.mark.parametrize("lattice,S,shells,key,kind", [("bcc", 4, 4, "ex_AlTiCrMo_V", 4), ("fcc", 5, 4, "ex_AlCrFeCoNi_V", 6),
                                                       ("fcc", 2, 4, "ex_FeNi_V", 6), ("bcc", 4, 6, "t02_V", 6)])
def test_full_size_128_cubed_properties(bw, golden, lattice, S, shells, key, kind):
    """BASELINE configs[1] at full size (128^3 bcc, 4 194 304 atoms) and the fcc / 6-shell lattices of the other configs at
    128^3: size-independent properties -- species counts conserved, occupancy pattern intact, sum of accepted dE == change
    of the exact (reference-order) total energy, exact and tree energies agree, trajectory reproducible."""
    n = 128
    V = golden[key][: S * S * shells]
    rng = np.random.default_rng(1)
    par = (np.arange(2 * n) & 1).astype(np.int8)
    if lattice == "bcc":
        mask = (par[None, None, :] == par[:, None, None]) & (par[None, :, None] == par[:, None, None])
    else:
        mask = ((par[None, None, :] + par[None, :, None] + par[:, None, None]) & 1) == 0
    N = int(mask.sum())
    g = np.zeros((2 * n,) * 3, dtype=np.int8)
    spec = np.repeat(np.arange(1, S + 1, dtype=np.int8), -(-N // S))[:N]; rng.shuffle(spec)
    g[mask] = spec
    finals = []
    for rep in range(2):
        dev = bw.Device(lattice, n, n, n, S, shells, V)
        assert dev.metropolis_plan()["use_box"] == kind
        dev.set_config(g)
        e0 = dev.total_energy(exact_order=True)[0]
        assert abs(dev.total_energy(exact_order=False)[0] - e0) < 1e-9 * abs(e0) + 1e-9
        att, acc, dE = dev.metropolis_run(1.0 / (1000.0 * bw.K_B_IN_RY), 4 * N, seed=5)
        g1 = dev.get_config()
        e1 = dev.total_energy(exact_order=True)[0]
        assert att[0] >= 4 * N and 0 < acc[0] < att[0]
        assert np.array_equal(np.bincount(g1.ravel(), minlength=S + 1), np.bincount(g.ravel(), minlength=S + 1))
        assert np.array_equal(g1 == 0, g == 0)
        # sum of accepted dE of the screened epoch kernels: fixed-point from a 23-bit table, < 2e-10 Ry per accepted swap
        # (the DECISIONS are exact; dE_mode 0 returns the sum to f64 rounding)
        assert abs((e1 - e0) - dE[0]) < 2e-10 * float(acc[0]), (e0, e1, dE[0])
        finals.append((g1, att[0], acc[0], e1))
        dev.close()
    assert np.array_equal(finals[0][0], finals[1][0]) and finals[0][1:] == finals[1][1:]


def test_random_config_on_device(bw, orc, golden):
    """brawl_cuda_random_config: start states for a batch of replicas generated on the device.  Exact
    species quotas per replica, sites only where the lattice has them, deterministic per (seed, offset, replica),
    independent across replicas, uniform over sites (chi-square on per-site species frequencies over 512 replicas),
    and a random-alloy energy that matches the oracle's host initial_setup ensemble."""
    V = golden["ex_AlCrFeCoNi_V"][: 5 * 5 * 4]
    R = 512
    dev = bw.Device("bcc", 4, 4, 4, 5, 4, V, n_replicas=R)
    counts = [26, 26, 26, 25, 25]
    dev.random_config(counts, 0, R, seed=123, offset=7)
    g = dev.get_config(0, R)
    osys = orc.System("bcc", 4, 4, 4, 5, 4, V)
    mask = osys.initial_setup(orc.MT(seed=1), *osys.quotas(conc=[0.2] * 5)) > 0
    assert np.all((g > 0) == mask[None])
    for r in range(R):
        assert np.array_equal(np.bincount(g[r].ravel(), minlength=6), [512 - 128] + counts)
    flat = g[:, mask]                                                     # [R][128] species 1..5
    assert len({row.tobytes() for row in flat}) == R                      # all replicas differ
    dev.random_config(counts, 0, R, seed=123, offset=7)                   # deterministic
    assert np.array_equal(dev.get_config(0, R), g)
    dev.random_config(counts, 5, 3, seed=123, offset=7)                   # a sub-range regenerates the same replicas
    assert np.array_equal(dev.get_config(0, R), g)
    dev.random_config(counts, 0, R, seed=123, offset=8)
    g2 = dev.get_config(0, R)
    assert not np.array_equal(g2, g)
    # uniformity: species frequency at each site ~ Binomial(R, count/128); chi-square over 128 sites x 5 species
    both = np.concatenate([flat, g2[:, mask]])                            # 1024 samples per site
    n = both.shape[0]
    chi = 0.0
    for s, c in enumerate(counts, start=1):
        exp = n * c / 128.0
        chi += (((both == s).sum(axis=0) - exp) ** 2 / exp).sum()
    dof = 128 * 4
    assert abs(chi - dof) < 6 * np.sqrt(2 * dof), chi                     # 6 sigma
    # pair correlations vanish: mean energy equals the host ensemble's within errors
    e_dev = dev.total_energy(0, R)
    mt = orc.MT(seed=99)
    e_host = []
    for _ in range(256):
        conc, cnt = osys.quotas(numbers=counts)
        e_host.append(osys.total_energy(osys.initial_setup(mt, conc, cnt)))
    e_host = np.array(e_host)
    se = np.sqrt(e_dev.var() / R + e_host.var() / e_host.size)
    assert abs(e_dev.mean() - e_host.mean()) < 5 * se, (e_dev.mean(), e_host.mean(), se)
    # errors: wrong multiset, bad range
    with pytest.raises(bw.BrawlCudaError):
        dev.random_config([26, 26, 26, 25, 24], 0, 1)
    with pytest.raises(bw.BrawlCudaError):
        dev.random_config(counts, R - 1, 2)
    # the 128^3 bench lattice in one call (4 194 304 sites: chunking, 64-bit keys)
    V4 = golden["ex_AlTiCrMo_V"][: 4 * 4 * 4]
    big = bw.Device("bcc", 128, 128, 128, 4, 4, V4)
    big.random_config([1048576] * 4, seed=5)
    gb = big.get_config()
    assert np.array_equal(np.bincount(gb.ravel(), minlength=5), [256 ** 3 - 4194304] + [1048576] * 4)
    sub = gb[gb > 0].astype(np.int64)
    assert abs((sub[:-1] == sub[1:]).mean() - 0.25) < 2e-3                # neighbours in site order uncorrelated


def test_store_state_occupancies(bw, orc, golden):
    """store_state (analytics.f90:43-64) on the device: occupancy counts accumulated over a short Metropolis run equal
    the counts accumulated from the configurations read back, in the reference's order(species, 1, x, y, z) layout."""
    V = golden["ex_AlTiCrMo_V"][: 4 * 4 * 4]
    for lattice, n in (("bcc", (4, 4, 4)), ("fcc", (3, 4, 5))):
        sysm = orc.System(lattice, *n, 4, 4, V)
        R = 3
        dev = bw.Device(lattice, *n, 4, 4, V, n_replicas=R)
        assert not dev.get_order(1).any()                                      # nothing stored yet
        na = dev.n_atoms
        dev.random_config([na // 4] * 4, 0, R, seed=3)
        want = np.zeros((R,) + dev.shape + (4,))
        beta = np.full(R, 1.0 / (800.0 * bw.K_B_IN_RY))
        for k in range(5):
            dev.metropolis_run(beta, 10 * na, seed=k)
            dev.store_state(0, R)
            g = dev.get_config(0, R)
            for s in range(4):
                want[..., s] += (g == s + 1)                                   # the reference's loop, vectorised
        for r in range(R):
            got = dev.get_order(r)
            assert got.shape == dev.shape + (4,) and np.array_equal(got, want[r]), (lattice, r)
            assert np.array_equal(got.sum(axis=-1), 5.0 * (dev.get_config(r) > 0))   # every site counted once per sample
        assert np.array_equal(dev.get_order(1, reset=True), want[1])
        assert not dev.get_order(1).any() and np.array_equal(dev.get_order(2), want[2])
        dev.store_state(1, 1)                                                  # a sub-range only touches its replicas
        assert dev.get_order(1).sum() == na and np.array_equal(dev.get_order(0), want[0])


def test_radial_counts_batch_matches_single(bw, orc, golden):
    """brawl_cuda_radial_counts_batch: all replicas in one launch == the per-replica entry == the oracle."""
    V = golden["ex_AlCrFeCoNi_V"][: 5 * 5 * 4]
    for lattice, n, R in (("bcc", (6, 5, 7), 7), ("fcc", (4, 4, 4), 33)):
        sysm = orc.System(lattice, *n, 5, 4, V)
        dev = bw.Device(lattice, *n, 5, 4, V, n_replicas=R)
        na = dev.n_atoms
        dev.random_config([na // 5 + (1 if s < na % 5 else 0) for s in range(5)], 0, R, seed=17)
        rho = dev.radial_densities_batch(4, 0, R)
        assert rho.shape == (R, 4, 5, 5)
        for r in range(R):
            assert np.array_equal(rho[r], dev.radial_densities(4, r))
        g = dev.get_config(R - 1)
        assert np.array_equal(rho[R - 1], sysm.radial_densities(g, 4, sysm.lattice_shells(g, 4)))
        assert np.array_equal(dev.radial_densities_batch(4, 2, 3), rho[2:5])
        with pytest.raises(bw.BrawlCudaError):
            dev.radial_densities_batch(4, R - 1, 2)


@pytest.mark.parametrize("lattice,shells,S,key,T,kind", [("bcc", 4, 4, "ex_AlTiCrMo_V", 1000.0, 4), ("bcc", 4, 5, "ex_AlCrFeCoNi_V", 800.0, 4),
                                                         ("bcc", 6, 4, "t02_V", 1000.0, 6), ("fcc", 4, 5, "ex_AlCrFeCoNi_V", 800.0, 6)])
def test_epoch_kernel_statistics_match_oracle(bw, orc, golden, lattice, shells, S, key, T, kind):
    """The headline kernel (epoch kernel: dense non-interacting sets, site energies cached over 4 steps, use_box == 4) and
    the byte-lattice epoch kernels of the other geometries (6-shell bcc, fcc: period-P classes, use_box == 6)
    against the oracle's sequential reference sampler, directly, at the bench temperature: energy per atom, heat capacity
    and the Warren-Cowley parameters alpha = 1 - rho/(Z c) (examples/01_metropolis_FeNi/02_simulated_annealing/
    01_plot_results.py:35-36) of shells 1 and 2.  32^3 cells (bcc 65 536 atoms, fcc 131 072).  Six GPU replicas are equilibrated for 4000
    sweeps; every equilibrated configuration then starts one oracle chain (own MT stream) AND the continuation of its GPU
    replica: 30 sweeps discarded, 120 sweeps sampled every 5.  Both samplers leave the Boltzmann distribution invariant,
    so the averages of a pair must agree within the blocked statistical errors (blocks of 4 samples = 20 sweeps, 6 per
    chain; pairing removes the slow replica-to-replica differences of the ordered state): the mean pair difference must
    be below 6 standard errors (5 sigma + 1 sigma slack, no other tolerance)."""
    import threading
    V = golden[key][: S * S * shells]
    n, R = 32, 6
    sysm = orc.System(lattice, n, n, n, S, shells, V)
    N = sysm.n_atoms
    beta = 1.0 / (T * bw.K_B_IN_RY)
    dev = bw.Device(lattice, n, n, n, S, shells, V, n_replicas=R)
    assert dev.metropolis_plan()["use_box"] == kind and (kind != 4 or dev.metropolis_plan()["trials_per_step"] == 960)
    starts = []
    for r in range(R):
        rng = np.random.default_rng(40 + r)
        spec = np.repeat(np.arange(1, S + 1, dtype=np.int8), -(-N // S))[:N]
        rng.shuffle(spec)
        g = np.zeros((2 * n,) * 3, dtype=np.int8)
        par = np.arange(2 * n) & 1
        if lattice == "bcc":
            g[(par[None, None, :] == par[:, None, None]) & (par[None, :, None] == par[:, None, None])] = spec
        else:
            g[((par[None, None, :] + par[None, :, None] + par[:, None, None]) & 1) == 0] = spec
        starts.append(g)
    dev.set_config(np.stack(starts))
    dev.metropolis_run(beta, 4000 * N, seed=77)
    eq = dev.get_config(0, R).copy()
    conc = np.bincount(starts[0].ravel(), minlength=S + 1)[1:] / float(N)
    Z = np.array([8.0, 6.0]) if lattice == "bcc" else np.array([12.0, 6.0])
    BS = 4                                                                      # samples per block

    def observables(e_series, rho_series):
        """one chain: samples of E/N [m], rho [m][2][S][S] -> per block (E/N, C_V per atom, alpha[2][S][S])"""
        e = np.asarray(e_series); rho = np.asarray(rho_series)
        nb = e.size // BS
        eb = e[: nb * BS].reshape(nb, BS)
        alpha = 1.0 - rho[: nb * BS].reshape(nb, BS, 2, S, S).mean(axis=1) / (Z[None, :, None, None] * conc[None, None, :, None])
        cv = (eb * N).var(axis=1, ddof=1) / (bw.K_B_IN_RY * T * T) / N        # (<E^2> - <E>^2) / (k_B T^2) per atom
        return eb.mean(axis=1), cv, alpha

    orc_out = [None] * R
    def chain(c):                                                               # threads: ctypes releases the GIL
        g = eq[c].copy()
        mt = orc.MT(seed=3100 + c)
        sysm.metropolis_trials(g, mt, beta, 30 * N)
        es, rs = [], []
        shells = sysm.lattice_shells(g, 3)
        for _ in range(24):
            sysm.metropolis_trials(g, mt, beta, 5 * N)
            es.append(sysm.total_energy(g) / N)
            rs.append(sysm.radial_densities(g, 3, shells)[1:3])
        orc_out[c] = observables(es, rs)
    ths = [threading.Thread(target=chain, args=(c,)) for c in range(R)]
    for t in ths:
        t.start()
    dev.metropolis_run(beta, 30 * N, seed=78)                                   # the GPU replicas meanwhile
    es, rs = [[] for _ in range(R)], [[] for _ in range(R)]
    for k in range(24):
        dev.metropolis_run(beta, 5 * N, seed=100 + k)
        e = dev.total_energy(0, R, exact_order=False) / N
        rho = dev.radial_densities_batch(3, 0, R)
        for r in range(R):
            es[r].append(e[r]); rs[r].append(rho[r, 1:3])
    gpu_out = [observables(es[r], rs[r]) for r in range(R)]
    for t in ths:
        t.join()

    worst, report = 0.0, []
    for i, name in enumerate(("E/N", "C_V", "alpha")):
        d = np.mean([gpu_out[c][i].mean(axis=0) - orc_out[c][i].mean(axis=0) for c in range(R)], axis=0)
        var = sum(gpu_out[c][i].var(axis=0, ddof=1) / gpu_out[c][i].shape[0] + orc_out[c][i].var(axis=0, ddof=1) / orc_out[c][i].shape[0]
                  for c in range(R)) / R ** 2
        se = np.sqrt(var)
        assert np.all(se > 0)
        z = np.abs(d) / se
        worst = max(worst, float(np.max(z)))
        report.append((name, np.max(np.abs(d)), np.max(se)))
        assert np.all(z < 6.0), (name, d, se, z)
    # the comparison has teeth: the errors are a small fraction of the signal
    e_mean = np.mean([o[0].mean() for o in orc_out])
    a_mean = np.mean([o[2].mean(axis=0) for o in orc_out], axis=0)
    assert report[0][2] < 5e-3 * abs(e_mean), (report, e_mean)
    assert np.max(np.abs(a_mean)) > 20 * report[2][2], (report, a_mean)        # SRO is resolved, not noise
    print("epoch kernel vs oracle, %s %d shells S=%d T=%g: <E>/N = %.7f, max |diff| / max se: %s, worst |z| = %.2f"
          % (lattice, shells, S, T, e_mean, ["%s %.2e / %.2e" % r for r in report], worst))


@pytest.mark.parametrize("n", [64, 96, 160])
def test_epoch_kernel_box_load_paths(bw, golden, n, monkeypatch):
    """The box of the epoch kernel arrives through the bulk-async copy engine (cp.async.bulk, whole x-rows of <= 128 bytes
    expanded in place) or, for wider lattices (n = 160) and with BRAWL_CUDA_NO_TMA set, through the LDG loop: same seed =>
    identical trajectory on both paths; composition conserved, accepted dE sum == energy change."""
    V = golden["ex_AlTiCrMo_V"][:64]
    beta = 1.0 / (1000.0 * bw.K_B_IN_RY)
    out = []
    for no_tma in (False, True):
        if no_tma:
            monkeypatch.setenv("BRAWL_CUDA_NO_TMA", "1")
        else:
            monkeypatch.delenv("BRAWL_CUDA_NO_TMA", raising=False)
        dev = bw.Device("bcc", n, n, n, 4, 4, V)
        N = dev.n_atoms
        dev.random_config([N // 4] * 4, seed=17)
        lat0 = dev.get_lattice()[0].copy()
        assert dev.metropolis_plan()["use_box"] == 4
        e0 = dev.total_energy(exact_order=False)[0]
        att, acc, dE = dev.metropolis_run(beta, 3 * N, seed=5)
        e1 = dev.total_energy(exact_order=False)[0]
        lat1 = dev.get_lattice()[0].copy()
        assert att[0] >= 3 * N and 0 < acc[0] < att[0]
        assert np.array_equal(np.bincount(lat0, minlength=4), np.bincount(lat1, minlength=4))
        assert abs((e1 - e0) - dE[0]) < 2e-10 * float(acc[0]) + 1e-9 * abs(e0), (e0, e1, dE[0])
        out.append((lat0, lat1, int(acc[0])))
        dev.close()
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1]) and out[0][2] == out[1][2]
