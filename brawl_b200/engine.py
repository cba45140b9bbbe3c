"""Host-side mirror of BraWl's operator table on top of the C ABI (include/brawl_cuda.h).

`RunParams` plays the role of the reference's `type(run_params)` (src/derived_types.f90:56-100):
it carries lattice, n_1..n_3, n_species, interaction_range and V_ex, and its bound operators keep
the reference's names and argument meaning --

    setup.full_energy(config)                       -> total_energy         (bw_hamiltonian.f90:58)
    setup.nbr_energy(config, b, i, j, k)            -> <lat>_energy_Nshells (1-based site, like Fortran)
    setup.pair_energy(config, idx1, idx2)           -> pair_energy          (bw_hamiltonian.f90:99)
    setup.mc_steps(config, beta, n, mt)             -> n x setup%mc_step    (metropolis.F90:751-891)

-- but every one of them runs on the GPU through libbrawl_cuda.so.  `config` is the reference's
array: numpy int8 of shape (2*n_3, 2*n_2, 2*n_1) (= Fortran config(1,:,:,:) memory order), species
1..S, 0 off-site.  Errors raise BrawlCudaError (the reference `stop`s).  No CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import BrawlCudaError, check

LATTICES = {"simple_cubic": 0, "bcc": 1, "fcc": 2}

# src/constants.f90:41-49 -- k_b_in_eV carries the reference's digit transposition
K_B_IN_RY = 8.167333262e-5 / 13.605693122990
RY_TO_EV = 13.605693122


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def rank_seed(seed, rank):
    """64-bit Philox key of a rank's Monte-Carlo kernels: the common seed for rank 0, a splitmix64 scramble of
    (seed, rank) otherwise -- distinct streams per rank (replica ids in the counters are handle-local)."""
    if rank == 0:
        return int(seed) & 0xFFFFFFFFFFFFFFFF
    z = (int(seed) + 0x9E3779B97F4A7C15 * int(rank)) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


class Device:
    """One handle = `n_replicas` lattices resident in HBM on one GPU (C ABI object)."""

    def __init__(self, lattice, n_1, n_2, n_3, n_species, n_shells, V_ex, device=0, n_replicas=1):
        self.L = _lib.load()
        lat = LATTICES[lattice] if isinstance(lattice, str) else int(lattice)
        V = np.ascontiguousarray(np.asarray(V_ex, dtype=np.float64).ravel()[: n_species * n_species * n_shells])
        if V.size != n_species * n_species * n_shells:
            raise BrawlCudaError("V_ex needs n_species^2 * n_shells entries")
        h = C.c_void_p()
        check(self.L.brawl_cuda_create(lat, n_1, n_2, n_3, n_species, n_shells, _p(V), device, n_replicas, C.byref(h)))
        self.h = h
        self.shape = (2 * n_3, 2 * n_2, 2 * n_1)
        self.n_replicas = n_replicas
        self.S = n_species
        na, gb, z, nr = C.c_int64(), C.c_int64(), C.c_int(), C.c_int()
        check(self.L.brawl_cuda_info(self.h, C.byref(na), C.byref(gb), C.byref(z), C.byref(nr)))
        self.n_atoms, self.grid_bytes, self.z_total = na.value, gb.value, z.value
        self._offset = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.brawl_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- plumbing ------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        check(self.L.brawl_cuda_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        check(self.L.brawl_cuda_synchronize(self.h))

    def _grids(self, config, n):
        g = np.ascontiguousarray(config, dtype=np.int8)
        if g.size != n * self.grid_bytes:
            raise BrawlCudaError("config has %d bytes, expected %d x %d" % (g.size, n, self.grid_bytes))
        return g

    def set_config(self, config, first_replica=0, n=None):
        n = (np.asarray(config).size // self.grid_bytes) if n is None else n
        g = self._grids(config, n)
        check(self.L.brawl_cuda_set_config(self.h, first_replica, n, _p(g)))

    def get_config(self, first_replica=0, n=1, out=None):
        g = np.empty((n,) + self.shape, dtype=np.int8) if out is None else out
        check(self.L.brawl_cuda_get_config(self.h, first_replica, n, _p(g)))
        return g[0] if (n == 1 and out is None) else g

    def set_lattice(self, sites, first_replica=0, n=None):
        """Compact configurations [n][n_atoms] uint8, species 0..S-1 in the device's site order (include/brawl_cuda.h)."""
        a = np.ascontiguousarray(sites, dtype=np.uint8)
        n = a.size // self.n_atoms if n is None else n
        if a.size != n * self.n_atoms:
            raise BrawlCudaError("set_lattice: expected %d x %d bytes" % (n, self.n_atoms))
        check(self.L.brawl_cuda_set_lattice(self.h, first_replica, n, _p(a)))

    def get_lattice(self, first_replica=0, n=1, out=None):
        a = np.empty((n, self.n_atoms), dtype=np.uint8) if out is None else out
        check(self.L.brawl_cuda_get_lattice(self.h, first_replica, n, _p(a)))
        return a

    def random_config(self, species_count, first_replica=0, n=1, seed=0x42726157, offset=0):
        """Independent uniformly random arrangements of the species multiset for replicas [first, first+n), generated
        on the device (initial_setup / WL re-randomisation for batches)."""
        cnt = np.ascontiguousarray(species_count, dtype=np.int64)
        if cnt.size != self.S:
            raise BrawlCudaError("species_count has %d entries, the system has %d species" % (cnt.size, self.S))
        check(self.L.brawl_cuda_random_config(self.h, first_replica, n, _p(cnt), seed, offset))

    def store_state(self, first_replica=0, n=1):
        """store_state (src/analytics.f90:43-64): add the current occupancies to the device-side site counts."""
        check(self.L.brawl_cuda_store_state(self.h, first_replica, n))

    def get_order(self, replica=0, reset=False):
        """Occupancy counts as float64 [2n3][2n2][2n1][S] = the reference's order(species, 1, x, y, z)."""
        out = np.empty(self.shape + (self.S,), dtype=np.float64)
        check(self.L.brawl_cuda_get_order(self.h, replica, _p(out), int(reset)))
        return out

    def copy_replica(self, src, dst):
        check(self.L.brawl_cuda_copy_replica(self.h, src, dst))

    # --- Hamiltonian ----------------------------------------------------------------------------
    def total_energy(self, first_replica=0, n=1, exact_order=True):
        e = np.zeros(n, dtype=np.float64)
        check(self.L.brawl_cuda_total_energy(self.h, first_replica, n, int(bool(exact_order)), _p(e)))
        return e

    def site_energies(self, replica=0):
        out = np.zeros(self.shape, dtype=np.float64)
        check(self.L.brawl_cuda_site_energies(self.h, replica, _p(out)))
        return out

    def nbr_energy(self, x, y, z, species=0, replica=0):
        """setup%nbr_energy for the 0-based grid site (x, y, z); species > 0 overrides the centre species."""
        e = C.c_double()
        check(self.L.brawl_cuda_nbr_energy(self.h, replica, int(x), int(y), int(z), int(species), C.byref(e)))
        return e.value

    def pair_dE(self, idx1, idx2, replica=0):
        i1 = np.ascontiguousarray(idx1, dtype=np.int32)
        i2 = np.ascontiguousarray(idx2, dtype=np.int32)
        if i1.shape != i2.shape:
            raise BrawlCudaError("idx1/idx2 shapes differ")
        out = np.zeros(i1.size, dtype=np.float64)
        check(self.L.brawl_cuda_pair_dE(self.h, replica, i1.size, _p(i1), _p(i2), _p(out)))
        return out

    # --- Metropolis -----------------------------------------------------------------------------
    def metropolis_replay(self, beta, n_trials, mt_state625, nbr_swap=False, replica=0, n_sample_steps=0):
        """mt_state625: numpy uint32[625] (mt[624] + mti), updated in place."""
        st = mt_state625
        assert st.dtype == np.uint32 and st.size == 625 and st.flags.c_contiguous
        acc = C.c_int64()
        if n_sample_steps:
            e = np.zeros(n_trials // n_sample_steps, dtype=np.float64)
            check(self.L.brawl_cuda_metropolis_replay_sampled(self.h, replica, beta, n_trials, n_sample_steps,
                                                              int(nbr_swap), _p(st), C.byref(acc), _p(e)))
            return acc.value, e
        check(self.L.brawl_cuda_metropolis_replay(self.h, replica, beta, n_trials, int(nbr_swap), _p(st), C.byref(acc)))
        return acc.value

    def _betas(self, beta):
        b = np.ascontiguousarray(np.broadcast_to(np.asarray(beta, dtype=np.float64), (self.n_replicas,)))
        return b

    def metropolis_run(self, beta, n_trials, seed=0x42726157, nbr_swap=False, offset=None):
        """Production run; returns (attempted[r], accepted[r], dE_sum[r])."""
        b = self._betas(beta)
        off = self._offset if offset is None else offset
        nxt = C.c_uint64()
        att = np.zeros(self.n_replicas, dtype=np.int64)
        acc = np.zeros(self.n_replicas, dtype=np.int64)
        dE = np.zeros(self.n_replicas, dtype=np.float64)
        check(self.L.brawl_cuda_metropolis_run(self.h, _p(b), n_trials, int(nbr_swap), seed, off, C.byref(nxt),
                                               _p(att), _p(acc), _p(dE)))
        self._offset = nxt.value
        return att, acc, dE

    def metropolis_enqueue(self, beta, n_trials, seed=0x42726157, nbr_swap=False, offset=None):
        """Asynchronous form: returns (attempts planned per replica, kernel launches)."""
        b = self._betas(beta)
        off = self._offset if offset is None else offset
        nxt, planned, nl = C.c_uint64(), C.c_int64(), C.c_int()
        check(self.L.brawl_cuda_metropolis_enqueue(self.h, _p(b), n_trials, int(nbr_swap), seed, off, C.byref(nxt),
                                                   C.byref(planned), C.byref(nl)))
        self._offset = nxt.value
        return planned.value, nl.value

    def metropolis_counters(self, reset=True):
        att = np.zeros(self.n_replicas, dtype=np.int64)
        acc = np.zeros(self.n_replicas, dtype=np.int64)
        dE = np.zeros(self.n_replicas, dtype=np.float64)
        check(self.L.brawl_cuda_metropolis_counters(self.h, int(reset), _p(att), _p(acc), _p(dE)))
        return att, acc, dE

    def metropolis_tune(self, box=(0, 0, 0), steps_per_phase=0):
        check(self.L.brawl_cuda_metropolis_tune(self.h, box[0], box[1], box[2], steps_per_phase))

    def metropolis_set_mode(self, dE_mode):
        """0: reference f64 association for every trial; 1: integer-count screening on the byte lattice with
        exact recomputation inside the guard band; 2 (default): the same screening on a word lattice with
        fixed-point dp4a dE and the dense decomposition where a word kernel is instantiated (else like 1).
        Screened and exact kernels on the same decomposition take identical decisions."""
        check(self.L.brawl_cuda_metropolis_set_mode(self.h, int(dE_mode)))

    def metropolis_set_layout(self, byte_layout_only):
        """True: never use the word-lattice kernels / dense decomposition (A/B comparisons); False: automatic."""
        check(self.L.brawl_cuda_metropolis_set_layout(self.h, int(byte_layout_only)))

    def metropolis_plan(self, nbr_swap=False):
        o = np.zeros(10, dtype=np.int32)
        check(self.L.brawl_cuda_metropolis_plan(self.h, int(nbr_swap), _p(o)))
        keys = ("use_box", "P", "margin", "box_x", "box_y", "box_z", "trials_per_step", "boxes_per_replica",
                "n_displacements", "steps_per_phase")
        d = dict(zip(keys, (int(v) for v in o)))
        d["n_orientations"] = (d["use_box"] >> 4) & 255
        d["warp_groups"] = 2 if (d["use_box"] >> 12) & 1 else 1
        d["use_box"] &= 15
        d["P"] = (d["P"] // 10000, (d["P"] // 100) % 100, d["P"] % 100)
        return d

    def metropolis_last_launches(self):
        n = C.c_int()
        check(self.L.brawl_cuda_metropolis_last_launches(self.h, C.byref(n)))
        return n.value

    # --- SRO --------------------------------------------------------------------------------------
    def radial_counts(self, wc_range, replica=0):
        cnt = np.zeros((wc_range, self.S, self.S), dtype=np.int64)
        sc = np.zeros(self.S, dtype=np.int64)
        check(self.L.brawl_cuda_radial_counts(self.h, replica, wc_range, _p(cnt), _p(sc)))
        return cnt, sc

    def radial_densities(self, wc_range, replica=0):
        """rho[l][j][i] = r_densities(i,j,l) of src/analytics.f90:293-404 (disk order of `rho data`)."""
        cnt, sc = self.radial_counts(wc_range, replica)
        return cnt / sc[None, None, :].astype(np.float64)

    def radial_densities_batch(self, wc_range, first_replica=0, n=1):
        """rho[r][l][j][i] for replicas [first, first+n) from one launch."""
        cnt = np.zeros((n, wc_range, self.S, self.S), dtype=np.int64)
        sc = np.zeros((n, self.S), dtype=np.int64)
        check(self.L.brawl_cuda_radial_counts_batch(self.h, first_replica, n, wc_range, _p(cnt), _p(sc)))
        return cnt / sc[:, None, None, :].astype(np.float64)

    # --- Wang-Landau --------------------------------------------------------------------------------
    def wl_sweeps_replay(self, lng, hist, bin_edges, win_lo, win_hi, wl_f, n_trials, mt_state625, nbr_swap=False,
                         replica=0):
        acc, ef = C.c_int64(), C.c_double()
        edges = np.ascontiguousarray(bin_edges, dtype=np.float64)
        check(self.L.brawl_cuda_wl_sweeps_replay(self.h, replica, _p(lng), _p(hist), _p(edges), lng.size, win_lo, win_hi,
                                                 wl_f, n_trials, int(nbr_swap), _p(mt_state625), C.byref(acc), C.byref(ef)))
        return acc.value, ef.value

    def wl_sweeps(self, lng, hist, bin_edges, win_lo, win_hi, wl_f, n_trials, seed=0x42726157, offset=0,
                  nbr_swap=False):
        """lng[W][bins], hist[W][stride] host arrays updated in place; returns (accepted[W], e_final[W])."""
        W, bins = lng.shape
        lo = np.ascontiguousarray(np.broadcast_to(win_lo, (W,)), dtype=np.int32)
        hi = np.ascontiguousarray(np.broadcast_to(win_hi, (W,)), dtype=np.int32)
        edges = np.ascontiguousarray(bin_edges, dtype=np.float64)
        acc = np.zeros(W, dtype=np.int64)
        ef = np.zeros(W, dtype=np.float64)
        check(self.L.brawl_cuda_wl_sweeps(self.h, W, _p(lng), _p(hist), 0, _p(edges), bins, _p(lo), _p(hi), hist.shape[1],
                                          wl_f, n_trials, int(nbr_swap), seed, offset, _p(acc), _p(ef)))
        return acc, ef

    def wl_enter_window(self, target, lo_e, hi_e, inv_two_sigma_sq, max_trials, seed=0x42726157, offset=0):
        n = len(target)
        t = np.ascontiguousarray(target, dtype=np.float64)
        lo = np.ascontiguousarray(lo_e, dtype=np.float64)
        hi = np.ascontiguousarray(hi_e, dtype=np.float64)
        e = np.zeros(n, dtype=np.float64)
        ent = np.zeros(n, dtype=np.int32)
        check(self.L.brawl_cuda_wl_enter_window(self.h, n, _p(t), _p(lo), _p(hi), inv_two_sigma_sq, int(max_trials), seed,
                                                offset, _p(e), _p(ent)))
        return e, ent

    def wl_enter_window_replay(self, e_start, target, lo_e, hi_e, two_sigma_sq, period, i_steps, max_iters, resume,
                               mt_state625, replica=0):
        """One stretch of enter_energy_window (wang-landau.F90:643-741) on the reference's MT stream; returns
        (status, e_running, i_steps, iterations begun).  status 1: window test passed on the running energy, 2: the
        lattice is due for re-randomisation, 0: max_iters reached (see include/brawl_cuda.h)."""
        isteps, e, st, it = C.c_int64(int(i_steps)), C.c_double(), C.c_int(), C.c_int64()
        check(self.L.brawl_cuda_wl_enter_window_replay(self.h, replica, e_start, target, lo_e, hi_e, two_sigma_sq,
                                                       int(period), C.byref(isteps), int(max_iters), int(resume),
                                                       _p(mt_state625), C.byref(e), C.byref(st), C.byref(it)))
        return st.value, e.value, isteps.value, it.value

    # device-resident Wang-Landau state and the collectives of the multi-GPU drivers (include/brawl_cuda.h)
    def wl_init(self, bins, bin_edges, walkers_per_window):
        edges = np.ascontiguousarray(bin_edges, dtype=np.float64)
        check(self.L.brawl_cuda_wl_init(self.h, int(bins), _p(edges), int(walkers_per_window)))
        self._wl_bins, self._wl_windows = int(bins), self.n_replicas // int(walkers_per_window)

    def wl_set_windows(self, win_lo, win_hi, zero_hist=True):
        lo = np.ascontiguousarray(np.broadcast_to(win_lo, (self.n_replicas,)), dtype=np.int32)
        hi = np.ascontiguousarray(np.broadcast_to(win_hi, (self.n_replicas,)), dtype=np.int32)
        check(self.L.brawl_cuda_wl_set_windows(self.h, _p(lo), _p(hi), int(zero_hist)))

    def wl_set_span(self, n_ranks):
        """The windows of this handle are shared with the same windows on all n_ranks ranks of its communicator: the window
        average of wl_iterate becomes local sum + ncclAllReduce (include/brawl_cuda.h)."""
        check(self.L.brawl_cuda_wl_set_span(self.h, int(n_ranks)))

    def wl_zero_hist(self):
        check(self.L.brawl_cuda_wl_zero_hist(self.h))

    def wl_set_lng(self, lng):
        a = np.ascontiguousarray(lng, dtype=np.float64)
        if a.size != self._wl_bins:
            raise BrawlCudaError("ln g has %d entries, the state has %d bins" % (a.size, self._wl_bins))
        check(self.L.brawl_cuda_wl_set_lng(self.h, _p(a)))

    def wl_get(self, what=0):
        """what = 0: ln g, 1: hist (entry i = bin win_lo + i) of every local window -> [n_windows][bins]"""
        out = np.zeros((self._wl_windows, self._wl_bins))
        check(self.L.brawl_cuda_wl_get(self.h, int(what), _p(out)))
        return out

    def wl_iterate(self, wl_f, n_trials, seed=0x42726157, offset=0, nbr_swap=False, want_accept=False):
        """sweeps + window average + flatness inputs on the device -> (energies[W], hist_min[windows], hist_mean[windows])"""
        e, mn, mean = np.zeros(self.n_replicas), np.zeros(self._wl_windows), np.zeros(self._wl_windows)
        acc = np.zeros(self.n_replicas, dtype=np.int64) if want_accept else None
        check(self.L.brawl_cuda_wl_iterate(self.h, wl_f, int(n_trials), int(nbr_swap), seed, offset, _p(e), _p(mn), _p(mean),
                                           _p(acc) if want_accept else None))
        return (e, mn, mean, acc) if want_accept else (e, mn, mean)

    def comm_unique_id(self):
        uid = np.zeros(128, dtype=np.uint8)
        check(self.L.brawl_cuda_comm_unique_id(_p(uid)))
        return uid

    def comm_create(self, n_ranks, rank, unique_id):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        check(self.L.brawl_cuda_comm_create(self.h, int(n_ranks), int(rank), _p(uid)))
        self._comm_ranks = int(n_ranks)

    def comm_destroy(self):
        check(self.L.brawl_cuda_comm_destroy(self.h))

    def comm_allgather(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        out = np.zeros((self._comm_ranks, a.size))
        check(self.L.brawl_cuda_comm_allgather(self.h, _p(a), a.size, _p(out)))
        return out

    def comm_allreduce(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64).copy()
        check(self.L.brawl_cuda_wl_allreduce(self.h, _p(a), a.size))
        return a

    def wl_allgather_lng(self, n_ranks):
        out = np.zeros((n_ranks * self._wl_windows, self._wl_bins))
        check(self.L.brawl_cuda_wl_allgather_lng(self.h, _p(out)))
        return out

    def exchange_replica(self, replica, peer):
        check(self.L.brawl_cuda_exchange_replica(self.h, int(replica), int(peer)))

    def exchange_replicas(self, replicas, peers):
        r = np.ascontiguousarray(replicas, dtype=np.int32)
        q = np.ascontiguousarray(peers, dtype=np.int32)
        check(self.L.brawl_cuda_exchange_replicas(self.h, r.size, _p(r), _p(q)))

    def copy_replicas_batch(self, src, dst):
        """replica dst[i] := replica src[i], all pairs in one launch (walker cloning of batched nested-sampling runs)"""
        s = np.ascontiguousarray(src, dtype=np.int32)
        d = np.ascontiguousarray(dst, dtype=np.int32)
        check(self.L.brawl_cuda_copy_replicas_batch(self.h, s.size, _p(s), _p(d)))

    def swap_replicas_batch(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.int32)
        b = np.ascontiguousarray(b, dtype=np.int32)
        check(self.L.brawl_cuda_swap_replicas_batch(self.h, a.size, _p(a), _p(b)))

    def swap_replicas(self, a, b):
        check(self.L.brawl_cuda_swap_replicas(self.h, a, b))

    def lattice_tensor(self, torch):
        """The device-resident compact lattices as a torch uint8 tensor [n_replicas][n_atoms] (no copy)."""
        ptr, nb = C.c_void_p(), C.c_int64()
        check(self.L.brawl_cuda_lattice_ptr(self.h, C.byref(ptr), C.byref(nb)))

        class _Arr:
            __cuda_array_interface__ = {"shape": (self.n_replicas, nb.value), "typestr": "|u1",
                                        "data": (ptr.value, False), "version": 2}
        return torch.as_tensor(_Arr(), device="cuda")

    # --- nested sampling ------------------------------------------------------------------------------
    def ns_walk_replay(self, energy, e_limit, n_steps, mt_state625, replica=0):
        e, acc = C.c_double(energy), C.c_int64()
        check(self.L.brawl_cuda_ns_walk_replay(self.h, replica, C.byref(e), e_limit, n_steps, _p(mt_state625), C.byref(acc)))
        return e.value, acc.value

    def ns_walk(self, walker_ids, energies, e_limit, n_steps, seed=0x42726157, offset=0):
        ids = np.ascontiguousarray(walker_ids, dtype=np.int32)
        e = np.ascontiguousarray(energies, dtype=np.float64).copy()
        lim = np.ascontiguousarray(np.broadcast_to(e_limit, ids.shape), dtype=np.float64)
        acc = np.zeros(ids.size, dtype=np.int64)
        check(self.L.brawl_cuda_ns_walk(self.h, ids.size, _p(ids), _p(e), _p(lim), n_steps, seed, offset, _p(acc)))
        return e, acc


class RunParams:
    """The reference's `setup` object with GPU-bound operators (src/derived_types.f90:56-100,
    bound as in initialise_function_pointers, src/initialise.F90:153-257)."""

    def __init__(self, lattice, n_1, n_2, n_3, n_species, interaction_range, V_ex, wc_range=2, device=0):
        self.lattice, self.n_1, self.n_2, self.n_3 = lattice, n_1, n_2, n_3
        self.n_species, self.interaction_range, self.wc_range = n_species, interaction_range, wc_range
        self.n_basis = 1
        self.dev = Device(lattice, n_1, n_2, n_3, n_species, interaction_range, V_ex, device=device, n_replicas=1)
        self.n_atoms = self.dev.n_atoms

    def full_energy(self, config):
        self.dev.set_config(config)
        return float(self.dev.total_energy(exact_order=True)[0])

    def nbr_energy(self, config, site_b, site_i, site_j, site_k):
        if site_b != 1:
            raise BrawlCudaError("n_basis is 1 for every supported lattice")
        self.dev.set_config(config)
        return self.dev.nbr_energy(site_i - 1, site_j - 1, site_k - 1)

    def _flat(self, idx):
        b, i, j, k = idx
        return ((k - 1) * 2 * self.n_2 + (j - 1)) * 2 * self.n_1 + (i - 1)

    def pair_energy_change(self, config, idx1, idx2):
        """pair_energy(after swap) - pair_energy(before) for 1-based (b,i,j,k) sites."""
        self.dev.set_config(config)
        return float(self.dev.pair_dE([self._flat(idx1)], [self._flat(idx2)])[0])

    def mc_steps(self, config, beta, n_trials, mt_state625, nbr_swap=False):
        """n_trials x setup%mc_step(config, beta) with the reference MT19937 stream; config is
        updated in place; returns the number of accepted trials."""
        self.dev.set_config(config)
        acc = self.dev.metropolis_replay(beta, n_trials, mt_state625, nbr_swap=nbr_swap)
        config[...] = self.dev.get_config()
        return acc

    def radial_densities(self, config, wc_range=None):
        self.dev.set_config(config)
        return self.dev.radial_densities(self.wc_range if wc_range is None else wc_range)
