"""ctypes front-end of the CPU oracle (oracle/brawl_oracle.c).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from brawl_b200/.

Grids are numpy int8 arrays of shape (2*n3, 2*n2, 2*n1) = [z][y][x] (x fastest), species
1..S, 0 = no site: byte-for-byte the reference's config(1,x,y,z) (src/shared_data.f90:30).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LATTICES = {"simple_cubic": 0, "sc": 0, "bcc": 1, "fcc": 2}

# src/constants.f90:41-49 (k_b_in_eV carries the reference's digit transposition)
K_B_IN_RY = 8.167333262e-5 / 13.605693122990
RY_TO_EV = 13.605693122


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "brawl_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_sys_create.restype = C.c_void_p
        L.orc_sys_create.argtypes = [C.c_int] * 6 + [C.c_void_p]
        L.orc_sys_destroy.argtypes = [C.c_void_p]
        L.orc_sys_grid_size.restype = C.c_long
        for f in ("orc_nbr_energy", "orc_total_energy", "orc_pair_energy", "orc_pair_dE", "orc_mt_genrand"):
            getattr(L, f).restype = C.c_double
        L.orc_mt_int32.restype = C.c_uint32
        L.orc_mt_static_seed.restype = C.c_uint32
        L.orc_metropolis_trials.restype = C.c_int64
        L.orc_wl_sweeps.restype = C.c_int64
        _lib = L
    return _lib


class MT(C.Structure):
    """MT19937 state; `.state625()` exports mt[624] + mti for the CUDA replay entry point."""
    _fields_ = [("mt", C.c_uint32 * 624), ("mti", C.c_int32)]

    def __init__(self, seed=None, rank=None):
        super().__init__()
        if rank is not None:
            seed = lib().orc_mt_static_seed(int(rank))
        if seed is not None:
            lib().orc_mt_init(C.byref(self), C.c_uint32(seed))

    def genrand(self):
        return lib().orc_mt_genrand(C.byref(self))

    def int32(self):
        return lib().orc_mt_int32(C.byref(self))

    def state625(self):
        a = np.empty(625, dtype=np.uint32)
        a[:624] = np.frombuffer(self.mt, dtype=np.uint32)
        a[624] = self.mti
        return a

    def load625(self, a):
        C.memmove(self.mt, np.ascontiguousarray(a[:624], dtype=np.uint32).ctypes.data, 624 * 4)
        self.mti = int(a[624])


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class System:
    def __init__(self, lattice, n1, n2, n3, n_species, n_shells, V):
        self.lattice = LATTICES[lattice] if isinstance(lattice, str) else int(lattice)
        self.n = (n1, n2, n3)
        self.S, self.n_shells = n_species, n_shells
        self.V = np.ascontiguousarray(np.asarray(V, dtype=np.float64).ravel()[: n_species * n_species * n_shells])
        assert self.V.size == n_species * n_species * n_shells
        self.h = lib().orc_sys_create(self.lattice, n1, n2, n3, n_species, n_shells, _p(self.V))
        if not self.h:
            raise ValueError("unsupported lattice / shells")
        self.shape = (2 * n3, 2 * n2, 2 * n1)
        self.n_atoms = lib().orc_sys_n_atoms(C.c_void_p(self.h))
        self.z_total = lib().orc_sys_z_total(C.c_void_p(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_sys_destroy(C.c_void_p(self.h))
            self.h = None

    @property
    def hp(self):
        return C.c_void_p(self.h)

    def offsets(self):
        out = np.zeros((self.z_total, 4), dtype=np.int32)
        lib().orc_sys_offsets(self.hp, _p(out))
        return out

    # --- configuration -------------------------------------------------------------------
    def quotas(self, conc=None, numbers=None):
        """HEAD's species_count / species_concentrations derivation (initialise.F90:468-506)."""
        c = np.zeros(self.S + 1, dtype=np.float64)
        if conc is not None:
            c[1:] = conc
        cnt = np.zeros(self.S, dtype=np.int64)
        num = None if numbers is None else np.ascontiguousarray(numbers, dtype=np.int64)
        rc = lib().orc_species_quotas(self.hp, _p(c), None if num is None else _p(num), _p(cnt))
        if rc:
            raise ValueError("species counts do not sum to the number of lattice sites")
        return c, cnt

    def initial_setup(self, mt, conc, count):
        c = np.ascontiguousarray(conc, dtype=np.float64)
        assert c.size == self.S + 1 and c[0] == 0.0
        cnt = np.ascontiguousarray(count, dtype=np.int64)
        g = np.zeros(self.shape, dtype=np.int8)
        if lib().orc_initial_setup(self.hp, _p(c), _p(cnt), C.byref(mt), _p(g)):
            raise ValueError("bad quotas")
        return g

    # --- energies ------------------------------------------------------------------------
    def total_energy(self, g):
        return lib().orc_total_energy(self.hp, _p(g))

    def nbr_energy(self, g, x, y, z):
        return lib().orc_nbr_energy(self.hp, _p(g), int(x), int(y), int(z))

    def site_energies(self, g):
        out = np.zeros(self.shape, dtype=np.float64)
        lib().orc_site_energies(self.hp, _p(g), _p(out))
        return out

    def pair_dE(self, g, idx1, idx2):
        i1 = np.ascontiguousarray(idx1, dtype=np.int32)
        i2 = np.ascontiguousarray(idx2, dtype=np.int32)
        out = np.zeros(i1.size, dtype=np.float64)
        gg = np.ascontiguousarray(g)
        lib().orc_pair_dE_batch(self.hp, _p(gg), C.c_long(i1.size), _p(i1), _p(i2), _p(out))
        return out

    # --- samplers ------------------------------------------------------------------------
    def mc_step(self, g, mt, beta, nbr_swap=False):
        return lib().orc_mc_step(self.hp, _p(g), C.byref(mt), C.c_double(beta), int(nbr_swap))

    def metropolis_trials(self, g, mt, beta, n_trials, nbr_swap=False):
        return lib().orc_metropolis_trials(self.hp, _p(g), C.byref(mt), C.c_double(beta),
                                           C.c_int64(n_trials), int(nbr_swap))

    def metropolis_sample(self, g, mt, temp, n_mc_steps, n_sample_steps, nbr_swap=False):
        n_sweeps = n_mc_steps // n_sample_steps
        e = np.zeros(n_sweeps, dtype=np.float64)
        out = np.zeros(3, dtype=np.float64)
        lib().orc_metropolis_sample(self.hp, _p(g), C.byref(mt), C.c_double(temp), C.c_double(K_B_IN_RY),
                                    C.c_int64(n_mc_steps), C.c_int64(n_sample_steps), int(nbr_swap),
                                    _p(e), _p(out))
        return e, out

    def wl_sweeps(self, g, mt, lng, hist, edges, win_lo, win_hi, wl_f, n_trials, nbr_swap=False):
        bins = lng.size
        assert edges.size == bins + 1 and hist.size == win_hi - win_lo + 1
        ef = C.c_double(0.0)
        acc = lib().orc_wl_sweeps(self.hp, _p(g), C.byref(mt), _p(lng), _p(hist), _p(edges), bins,
                                  int(win_lo), int(win_hi), C.c_double(wl_f), C.c_int64(n_trials),
                                  int(nbr_swap), C.byref(ef))
        return acc, ef.value

    def wl_enter_energy_window(self, g, mt, conc, count, min_e, max_e, energy_min, energy_max, max_iters):
        """enter_energy_window (src/wang-landau.F90:643-741), one walker -> (entered, e_final, loop iterations)."""
        c = np.ascontiguousarray(conc, dtype=np.float64)
        cnt = np.ascontiguousarray(count, dtype=np.int64)
        ef, it = C.c_double(0.0), C.c_int64(0)
        ok = lib().orc_wl_enter_energy_window(self.hp, _p(g), C.byref(mt), _p(c), _p(cnt), C.c_double(min_e), C.c_double(max_e),
                                              C.c_float(energy_min), C.c_float(energy_max), C.c_int64(max_iters),
                                              C.byref(ef), C.byref(it))
        return bool(ok), ef.value, it.value

    def wl_bin_edges(self, energy_min, energy_max, bins):
        """create_energy_bins (src/wang-landau.F90:969-986); energies in meV/atom -> Ry/cell."""
        energy_to_ry = self.n_atoms / (RY_TO_EV * 1000)
        # real(bins) is single precision; exact for integer bins
        bin_width = (energy_max - energy_min) / float(np.float32(bins)) * energy_to_ry
        return np.array([energy_min * energy_to_ry + i * bin_width for i in range(bins + 1)], dtype=np.float64)

    def nested_sampling(self, mt, conc, count, K, n_steps, n_iter):
        c = np.ascontiguousarray(conc, dtype=np.float64)
        cnt = np.ascontiguousarray(count, dtype=np.int64)
        walkers = np.zeros((K,) + self.shape, dtype=np.int8)
        energies = np.zeros(K, dtype=np.float64)
        culled = np.zeros(n_iter, dtype=np.float64)
        if lib().orc_nested_sampling(self.hp, C.byref(mt), _p(c), _p(cnt), K, n_steps, n_iter,
                                     _p(walkers), _p(energies), _p(culled)):
            raise ValueError("bad quotas")
        return culled, energies, walkers

    # --- observables ---------------------------------------------------------------------
    def lattice_shells(self, g, wc_range):
        out = np.zeros(wc_range, dtype=np.float64)
        lib().orc_lattice_shells(self.hp, _p(g), wc_range, _p(out))
        return out

    def radial_densities(self, g, wc_range, shells):
        """Returns rho[l][j][i] (disk order of `rho data`) = Fortran r_densities(i,j,l)."""
        out = np.zeros((wc_range, self.S, self.S), dtype=np.float64)
        sh = np.ascontiguousarray(shells, dtype=np.float64)
        lib().orc_radial_densities(self.hp, _p(g), wc_range, _p(sh), _p(out))
        return out

    def radial_counts(self, g, wc_range):
        cnt = np.zeros((wc_range, self.S, self.S), dtype=np.int64)
        sc = np.zeros(self.S, dtype=np.int64)
        lib().orc_radial_counts(self.hp, _p(g), wc_range, _p(cnt), _p(sc))
        return cnt, sc


def ref_mt_lib():
    """The reference's own mt19937ar.c compiled by oracle/Makefile into oracle/_ref/ (or None)."""
    so = os.path.join(HERE, "_ref", "libmt19937ar.so")
    if not os.path.exists(so):
        return None
    L = C.CDLL(so)
    L.genrand.restype = C.c_double
    L.genrand_int32.restype = C.c_ulong
    L.init_genrand.argtypes = [C.c_ulong]
    L.f90_init_genrand.restype = C.c_ulong
    L.f90_init_genrand.argtypes = [C.c_int, C.c_int, C.c_ulong]
    return L


def wl_mean_energy(lng, edges, bins, bin_width):
    """compute_mean_energy (src/wang-landau.F90:457-477) -> [300][2]."""
    lng = np.ascontiguousarray(lng, dtype=np.float64)
    edges = np.ascontiguousarray(edges, dtype=np.float64)
    out = np.zeros((300, 2))
    lib().orc_wl_mean_energy(lng.ctypes.data_as(C.c_void_p), edges.ctypes.data_as(C.c_void_p), C.c_int(bins),
                             C.c_double(bin_width), C.c_double(K_B_IN_RY), out.ctypes.data_as(C.c_void_p))
    return out


def wl_window_optimise(it, intervals, mc_steps, diffusion_prev, bins):
    """mpi_window_optimise rank-0 arithmetic (src/wang-landau.F90:1224-1311) -> (intervals, diffusion_prev)."""
    iv = np.ascontiguousarray(intervals, dtype=np.int64).copy()
    mc = np.ascontiguousarray(mc_steps, dtype=np.float64)
    prev = np.ascontiguousarray(diffusion_prev, dtype=np.float64).copy()
    lib().orc_wl_window_optimise(C.c_int(it), C.c_int(iv.shape[0]), iv.ctypes.data_as(C.c_void_p),
                                 mc.ctypes.data_as(C.c_void_p), prev.ctypes.data_as(C.c_void_p), C.c_int(bins))
    return iv, prev


def wl_dos_combine(lng_windows, window_indices):
    """dos_combine (src/wang-landau.F90:1147-1194) -> combined ln g [bins]."""
    lng = np.ascontiguousarray(lng_windows, dtype=np.float64)
    win = np.ascontiguousarray(window_indices, dtype=np.int64)
    out = np.zeros(lng.shape[1])
    lib().orc_wl_dos_combine(lng.ctypes.data_as(C.c_void_p), win.ctypes.data_as(C.c_void_p), C.c_int(lng.shape[0]),
                             C.c_int(lng.shape[1]), out.ctypes.data_as(C.c_void_p))
    return out


def wl_replica_exchange(energies, lng_ranks, window_indices, num_walkers, edges, mts):
    """replica_exchange (src/wang-landau.F90:1392-1519) for W*num_walkers ranks with their own MT streams
    (`mts`: ctypes array of MT, advanced in place) -> [(lower rank, upper rank), ...] of the exchanges made."""
    e = np.ascontiguousarray(energies, dtype=np.float64)
    lng = np.ascontiguousarray(lng_ranks, dtype=np.float64)
    win = np.ascontiguousarray(window_indices, dtype=np.int64)
    ed = np.ascontiguousarray(edges, dtype=np.float64)
    pairs = np.zeros((e.size, 2), dtype=np.int32)
    lib().orc_wl_replica_exchange.restype = C.c_int
    n = lib().orc_wl_replica_exchange(e.ctypes.data_as(C.c_void_p), lng.ctypes.data_as(C.c_void_p), win.ctypes.data_as(C.c_void_p),
                                      C.c_int(win.shape[0]), C.c_int(num_walkers), ed.ctypes.data_as(C.c_void_p),
                                      C.c_int(ed.size - 1), C.byref(mts), pairs.ctypes.data_as(C.c_void_p))
    return [tuple(int(v) for v in row) for row in pairs[:n]]
