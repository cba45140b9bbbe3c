"""Wang-Landau driver on top of the C ABI -- host-side mirror of the reference's wl_main
(src/wang-landau.F90:101-314) with the trial loop (`sweeps`, :539-626), the window entry
(`enter_energy_window`, :643-741) and the configuration exchange on the GPU.

Units of work are (window, walker) pairs.  Placement: every window lives entirely on one GPU
(rank r owns windows r*W/P .. (r+1)*W/P-1, `walkers` walkers each), so the per-`sweeps` average of
ln g / hist over a window's walkers (the MPI_Allreduce + "/num_walkers" at :628-631) never crosses
GPUs.  Cross-GPU traffic is only what the reference also has once per outer iteration / f-stage:
an all-gather of walker energies (replica_exchange, :1392-1501), point-to-point configuration
swaps between adjacent windows on different GPUs, the converged flags (:230) and the all-gather of
the W x bins ln g table for dos_combine (:1147-1194) -- all through torch.distributed (NCCL on
GPUs; gloo in the CPU tests of the host logic).

Dynamic window resizing (mpi_window_optimise, :1211-1328) follows the reference's `performance` switch: 0/1 resize
after pre-sampling and after every f-stage, 2/3 after pre-sampling only, 4 never (WLParams' default here; the
reference's file default is 0 and `from_file` keeps that).  After a resize the reference reloads a stored
configuration of a random bin of the new window (load_window_config, :1541-1554, from the bins x grid store that
energy_explore/merge_configs fill); here the walkers are steered from their current configurations into the new
window on the GPU by the enter_energy_window kernel instead -- no bins x grid store, same post-condition (every
walker inside its window).  compute_mean_energy (:457-477) is evaluated after every stage (`mean_energy`).
rho(E) (:574-592, save_rho_E :346-381) is sampled with the SRO kernel when `wc_range` > 0 (`_sample_rho`).
The pure functions below restate the reference's integer/f64 host arithmetic exactly.
"""
import math

import numpy as np

from .engine import Device, RY_TO_EV, K_B_IN_RY, BrawlCudaError, rank_seed


def _f32(v):
    return float(np.float32(v))


class WLParams:
    """wl_params (src/derived_types.f90:322-350); reals are single precision in the reference, so
    e.g. wl_f = 0.05 is really 0.05000000074505806."""

    def __init__(self, mc_sweeps=100, bins=512, num_windows=4, bin_overlap=0.25, tolerance=5e-5, flatness=0.9,
                 wl_f=0.05, energy_min=-96.0, energy_max=0.0, radial_samples=8, performance=4, nbr_swap=False):
        self.mc_sweeps, self.bins, self.num_windows = int(mc_sweeps), int(bins), int(num_windows)
        self.bin_overlap, self.tolerance, self.flatness = _f32(bin_overlap), _f32(tolerance), _f32(flatness)
        self.wl_f, self.energy_min, self.energy_max = _f32(wl_f), _f32(energy_min), _f32(energy_max)
        self.radial_samples, self.performance, self.nbr_swap = int(radial_samples), int(performance), bool(nbr_swap)

    @classmethod
    def from_file(cls, path):
        """read_wl_file (src/io.f90:951-1086): key=value lines, '#' comments, unknown keys ignored."""
        kw = {"performance": 0}                              # io.f90:981
        conv = dict(mc_sweeps=int, bins=int, num_windows=int, bin_overlap=float, tolerance=float, flatness=float,
                    wl_f=float, energy_min=float, energy_max=float, radial_samples=int, performance=int)
        for line in open(path):
            if "=" not in line or line.lstrip().startswith("#"):
                continue
            k, v = line.split("=", 1)
            k, v = k.rstrip(), v.split("#")[0].strip()     # a leading blank breaks a key in the reference too
            if k in conv:
                kw[k] = conv[k](v)
            elif k == "nbr_swap":
                kw[k] = v.strip(".").upper().startswith("T")
        return cls(**kw)


# --- pure host arithmetic (1-based inclusive bin indices, as in the reference) ---------------------
def divide_range(bins, num_windows):
    """divide_range (:855-891) with power = 1."""
    W = num_windows
    iv = np.zeros((W, 2), dtype=np.int64)
    iv[0, 0] = 1
    iv[W - 1, 1] = bins
    factor = (float(bins) - 1.0) / ((W + 1.0) ** 1 - 1.0)
    for i in range(2, W + 1):
        iv[i - 2, 1] = int(math.floor(factor * ((i - 1) ** 1) + 1))
        iv[i - 1, 0] = iv[i - 2, 1] + 1
    return iv


def create_overlap(intervals, bin_overlap):
    """create_overlap (:934-955): every window but the first is extended downwards; note the last
    window's width is computed without the +1 (reference quirk)."""
    idx = intervals.copy()
    W = idx.shape[0]
    if W > 1:
        for i in range(2, W):
            b = idx[i - 2, 1] - idx[i - 2, 0] + 1
            idx[i - 1, 0] = int(intervals[i - 1, 0] - max(math.ceil(np.float32(bin_overlap) * np.float32(b)), 2))
            idx[i - 1, 1] = intervals[i - 1, 1]
        b = idx[W - 2, 1] - idx[W - 2, 0]
        idx[W - 1, 0] = int(intervals[W - 1, 0] - max(math.ceil(np.float32(bin_overlap) * np.float32(b)), 2))
        idx[W - 1, 1] = intervals[W - 1, 1]
    return idx


def create_energy_bins(n_atoms, energy_min, energy_max, bins):
    """create_energy_bins (:969-986): meV/atom -> Ry/cell."""
    energy_to_ry = n_atoms / (RY_TO_EV * 1000)
    width = (energy_max - energy_min) / float(np.float32(bins)) * energy_to_ry
    return np.array([energy_min * energy_to_ry + i * width for i in range(bins + 1)], dtype=np.float64)


def bin_index(e, edges, bins):
    """bin_index (:515-523); int() truncates toward zero like Fortran INT."""
    return int(((e - edges[0]) / (edges[bins] - edges[0])) * float(bins)) + 1


def dos_combine(lng_windows, window_indices):
    """dos_combine (:1147-1194): stitch window i onto the combined curve at the overlap bin where
    the slopes agree best, then subtract the minimum.  lng_windows[W][bins] (already window-averaged)."""
    W, bins = lng_windows.shape
    comb = lng_windows[0].copy()
    beta_index = 0
    for i in range(2, W + 1):
        buf = lng_windows[i - 1]
        start, end = int(window_indices[i - 1, 0]), int(window_indices[i - 1, 1])
        beta_diff = np.finfo(np.float64).max
        for j in range(0, int(window_indices[i - 2, 1] - window_indices[i - 1, 0] - 1) + 1):
            b_orig = comb[start + j] - comb[start + j - 1]            # (start+j+1) - (start+j), 1-based
            b_merge = buf[start + j] - buf[start + j - 1]
            if abs(b_orig - b_merge) < beta_diff:
                beta_diff = abs(b_orig - b_merge)
                beta_index = start + j + 1
        for j in range(beta_index, end + 1):         # same association as the reference: (buf + comb(bi)) - buf(bi)
            comb[j - 1] = buf[j - 1] + comb[beta_index - 1] - buf[beta_index - 1]
    return comb - comb.min()


def _seq_sum(v):
    """Fortran SUM of a short f64 array as gfortran evaluates it without -ffast-math: left to right
    (Python's built-in sum() is compensated since 3.12 and numpy's is pairwise, so neither is used;
    np.add.accumulate IS the sequential left-to-right recurrence, 0.0 + v[0] + v[1] + ...)."""
    v = np.asarray(v, dtype=np.float64).ravel()
    return float(np.add.accumulate(v)[-1]) if v.size else 0.0


def energy_bin_width(n_atoms, energy_min, energy_max, bins):
    """bin_width as create_energy_bins leaves it in the module variable (:980)."""
    return (energy_max - energy_min) / float(np.float32(bins)) * (n_atoms / (RY_TO_EV * 1000))


def compute_mean_energy(lng, edges, bins, bin_width):
    """compute_mean_energy (:457-477): canonical mean energy from ln g at T = 10, 20, ... 3000 K.
    Returns mean_energy[300][2] = (<E>(T) in Ry per cell, beta)."""
    lng = np.asarray(lng, dtype=np.float64)
    buf = lng - lng.max()
    centre = np.asarray(edges[:bins], dtype=np.float64) + 0.5 * bin_width
    out = np.zeros((300, 2))
    for itemp in range(1, 301):
        beta = 1.0 / (K_B_IN_RY * itemp * 10.0)
        prob = buf[:bins] - beta * centre
        prob = np.exp(prob - prob.max())
        prob = prob / _seq_sum(prob)
        out[itemp - 1, 0] = _seq_sum(centre * prob)
        out[itemp - 1, 1] = beta
    return out


def sort_descending(a):
    """sort_descending (:1686-1701): the reference's exchange sort of an index vector (1-based indices
    returned); ties keep the order this particular algorithm produces, which decides where the
    left-over bins go in window_optimise."""
    n = len(a)
    idx = list(range(1, n + 1))
    for i in range(n - 1):
        for j in range(i + 1, n):
            if a[idx[i] - 1] < a[idx[j] - 1]:
                idx[i], idx[j] = idx[j], idx[i]
    return idx


def _nint(x):
    """Fortran NINT: round half away from zero."""
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def window_optimise(it, window_intervals, wl_mc_steps, diffusion_prev, bins):
    """The rank-0 arithmetic of mpi_window_optimise (:1211-1311): resize the windows so that windows which
    needed more trials per bin to become flat get fewer bins.  `it` is the argument the reference passes (0 after
    pre-sampling, iter-1 in the main loop); wl_mc_steps[W] = trials each window spent before it converged (summed
    over its walkers); diffusion_prev[W] = the previous blend (1/W initially, :1079).  Returns
    (new window_intervals [W][2], 1-based inclusive; new diffusion_prev).  The caller then applies create_overlap
    (mpi_arrays, :1351-1379).  Reference quirks kept: the per-window width used for the weights is
    ABS(first-last+1) = width-2; the floor w_min = 0.02; weights_log is computed but never used."""
    W = len(wl_mc_steps)
    iv = np.array(window_intervals, dtype=np.int64).copy()
    if W < 2:
        return iv, np.array(diffusion_prev, dtype=np.float64).copy()
    alpha = 0.8 * (0.8 ** (it - 1))
    if it == 0:
        alpha = 1.0
    w_min = 0.02
    prev = [float(x) for x in diffusion_prev]
    w_mc = []
    for i in range(W):
        first, last = int(iv[i, 0]), int(iv[i, 1])
        with np.errstate(divide="ignore", invalid="ignore"):          # width-2 window: x/0 = Inf, 1/Inf = 0 as in IEEE Fortran
            w_mc.append(float(np.float64(1.0) / (np.float64(wl_mc_steps[i]) / np.float64(np.float32(abs(first - last + 1))))))
    s = _seq_sum(w_mc)
    w_mc = [x / s for x in w_mc]
    frac = [alpha * a + (1.0 - alpha) * b for a, b in zip(w_mc, prev)]
    s = _seq_sum(frac)
    frac = [x / s for x in frac]
    new_prev = np.array(frac, dtype=np.float64)
    frac = [max(x, w_min) for x in frac]
    free = [x > w_min for x in frac]
    if abs(_seq_sum(frac) - 1.0) > 1.0e-12:
        rem = 1.0 - _seq_sum(frac)
        if any(free):
            sum_free = _seq_sum([x for x, f in zip(frac, free) if f])
            if sum_free > 0.0:
                scale = rem / sum_free
                frac = [x + x * scale if f else x for x, f in zip(frac, free)]
    s = _seq_sum(frac)
    frac = [x / s for x in frac]
    nb = [_nint(float(np.float32(bins)) * x) for x in frac]
    min_bins = max(int(w_min * bins), 2)
    nb = [max(b, min_bins) for b in nb]
    if sum(nb) != bins:
        idx = sort_descending(nb)
        diff = bins - sum(nb)
        i = 1
        guard = 0
        while diff != 0:
            j = idx[(i - 1) % W] - 1
            if diff > 0:
                nb[j] += 1
                diff -= 1
            elif nb[j] > min_bins:
                nb[j] -= 1
                diff += 1
            i += 1
            guard += 1
            if guard > 4 * W * bins:
                raise BrawlCudaError("window_optimise: %d bins cannot hold %d windows of >= %d bins" % (bins, W, min_bins))
    iv[0, 1] = nb[0]
    for i in range(1, W):
        iv[i, 0] = iv[i - 1, 1] + 1
        iv[i, 1] = iv[i, 0] + nb[i] - 1
    iv[W - 1, 0] = iv[W - 2, 1] + 1
    iv[W - 1, 1] = bins
    return iv, new_prev


def overlap_location(ibin, q, window_indices):
    """Which overlap region (1-based index of its lower window, 0 = none) a walker of window q
    (1-based) with energy bin ibin sits in (replica_exchange, :1409-1427)."""
    W = window_indices.shape[0]
    lower = q > 1 and (ibin < window_indices[q - 2, 1] + 1) and (ibin > window_indices[q - 1, 0] - 1)
    upper = q < W and (ibin > window_indices[q, 0] - 1) and (ibin < window_indices[q - 1, 1] + 1)
    return q if upper else (q - 1 if lower else 0)


def plan_replica_exchange(energies, lng_windows, window_indices, walkers, edges, rng):
    """Pair walkers of adjacent windows that sit in the same overlap region (random order, :1441-1468)
    and decide each exchange with the lower walker's ln g: u < exp(lng(ibin) - lng(jbin)) (:1482).
    energies[W*walkers] in global walker order (window-major).  Deterministic given `rng`, so every
    rank computes the same plan.  Returns [(walker_a, walker_b), ...] of accepted exchanges."""
    W = window_indices.shape[0]
    bins = edges.size - 1
    # bin_index and overlap_location of every walker at once (same f64 arithmetic, element by element)
    e = np.asarray(energies, dtype=np.float64)
    ib = (np.trunc(((e - edges[0]) / (edges[bins] - edges[0])) * float(bins))).astype(np.int64) + 1
    q = np.arange(W * walkers) // walkers + 1
    wi = np.asarray(window_indices, dtype=np.int64)
    lo_prev_hi = wi[np.maximum(q - 2, 0), 1]; lo_own = wi[q - 1, 0]; hi_own = wi[q - 1, 1]; up_next_lo = wi[np.minimum(q, W - 1), 0]
    lower_reg = (q > 1) & (ib < lo_prev_hi + 1) & (ib > lo_own - 1)
    upper_reg = (q < W) & (ib > up_next_lo - 1) & (ib < hi_own + 1)
    loc = np.where(upper_reg, q, np.where(lower_reg, q - 1, 0))
    swaps = []
    for i in range(1, W):
        lower = [g for g in range((i - 1) * walkers, i * walkers)]
        upper = [g for g in range(i * walkers, (i + 1) * walkers)]
        rng.shuffle(lower)
        rng.shuffle(upper)
        # every lower walker of the region takes the first unused upper walker of the region: the two filtered orders, zipped
        # (bins / regions are NOT refreshed after an exchange -- the reference evaluates them once per call (:1405-1431); a
        # walker sits in at most one region, so it is paired at most once)
        lower_c = [a for a in lower if loc[a] == i]
        if not lower_c:
            continue
        upper_c = [b for b in upper if loc[b] == i]
        lng = lng_windows[i - 1]
        for a, b in zip(lower_c, upper_c):
            if rng.random() < math.exp(min(0.0, lng[ib[a] - 1] - lng[ib[b] - 1])):
                swaps.append((a, b))
    return swaps


def random_configuration(lattice, n_1, n_2, n_3, counts, rng):
    """A uniformly random arrangement of the species multiset on the lattice sites (the distribution
    initial_setup samples, src/initialise.F90:434-617), reference grid layout."""
    z, y, x = np.meshgrid(np.arange(2 * n_3), np.arange(2 * n_2), np.arange(2 * n_1), indexing="ij")
    if lattice == "bcc":
        mask = ((x & 1) == (z & 1)) & ((y & 1) == (z & 1))
    elif lattice == "fcc":
        mask = ((x + y + z) & 1) == 0
    else:
        mask = np.ones_like(x, dtype=bool)
    spec = np.concatenate([np.full(int(c), s + 1, dtype=np.int8) for s, c in enumerate(counts)])
    if spec.size != int(mask.sum()):
        raise BrawlCudaError("species counts do not sum to the number of lattice sites")
    rng.shuffle(spec)
    g = np.zeros(mask.shape, dtype=np.int8)
    g[mask] = spec
    return g


from .netcdf3 import ncdf_writer_1d, ncdf_radial_density_writer_across_energy  # noqa: E402


class _Comm:
    """torch.distributed plumbing (world size 1 needs no torch at all): gloo in the CPU tests, NCCL on GPUs."""

    def __init__(self, rank=0, world=1, device=None):
        self.rank, self.world, self.device = rank, world, device
        if world > 1:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist

    def all_gather(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if self.world == 1:
            return a[None]
        t = self.torch.from_numpy(a.copy())
        if self.device is not None:
            t = t.to(self.device)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return np.stack([o.cpu().numpy() for o in out])

    def all_sum(self, v):
        return float(self.all_gather(np.array([float(v)])).sum())


class _AbiComm:
    """The same collectives through the C ABI's own NCCL communicator (brawl_cuda_comm_*, include/brawl_cuda.h) -- the
    calls a Fortran wl_main binds in place of its MPI ones.  `unique_id`: the 128 bytes rank 0 got from
    Device.comm_unique_id(), already distributed by the caller."""
    abi = True

    def __init__(self, dev, rank, world, unique_id):
        self.dev, self.rank, self.world = dev, rank, world
        if world > 1:
            dev.comm_create(world, rank, unique_id)

    def all_gather(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if self.world == 1:
            return a[None]
        return self.dev.comm_allgather(a.ravel()).reshape((self.world,) + a.shape)

    def all_sum(self, v):
        if self.world == 1:
            return float(v)
        return float(self.dev.comm_allreduce(np.array([float(v)]))[0])


class WangLandau:
    """wl_main for the windows owned by this rank.  `dev` is created here: one handle with (windows_per_rank * walkers)
    replicas; ln g and the histograms of its walkers stay on the device (brawl_cuda_wl_init / wl_iterate), the host sees
    8 bytes per walker and 16 per window per `sweeps` call.  comm = "torch" (torch.distributed: NCCL on GPUs, gloo in
    the CPU tests) or "abi" (the C ABI's NCCL communicator; needs `unique_id`)."""

    def __init__(self, lattice, n_1, n_2, n_3, n_species, n_shells, V_ex, counts, params, walkers=8, device=0,
                 rank=0, world=1, seed=0x42726157, torch_device=None, wc_range=0, comm="torch", unique_id=None, span=False):
        """span = False: the windows are sharded over the GPUs (num_windows divisible by world), every window lives on one
        GPU.  span = True: EVERY rank holds `walkers` walkers of EVERY window (walkers * world per window in total) and the
        window average of each `sweeps` call is a cross-GPU all-reduce (brawl_cuda_wl_set_span) -- for runs with fewer
        windows than GPUs, e.g. the reference's one-window performance/wl_input.inp."""
        self.p, self.walkers, self.rank, self.world, self.seed = params, walkers, rank, world, seed
        W = params.num_windows
        self.span = bool(span) and world > 1
        if not self.span and W % world:
            raise BrawlCudaError("num_windows must be divisible by the number of GPUs (or use span=True: windows shared by all GPUs)")
        self.w_local = W if self.span else W // world
        self.first_window = 0 if self.span else rank * self.w_local                 # 0-based
        self.n_local = self.w_local * walkers
        self.walkers_total = walkers * (world if self.span else 1)                  # walkers of one window over all ranks
        # window-major order of the global walker list: wm[j] = rank-major id (rank * n_local + local) of the j-th walker
        # when the walkers are sorted by window (identity when the windows are sharded in order)
        if self.span:
            self.wm = np.array([r * self.n_local + q * walkers + k for q in range(W) for r in range(world) for k in range(walkers)])
        else:
            self.wm = np.arange(W * walkers)
        self.dev = Device(lattice, n_1, n_2, n_3, n_species, n_shells, V_ex, device=device, n_replicas=self.n_local)
        self.lattice, self.n, self.counts = lattice, (n_1, n_2, n_3), counts
        self.n_atoms = self.dev.n_atoms
        self.intervals = divide_range(params.bins, W)
        self.window_indices = create_overlap(self.intervals, params.bin_overlap)
        self.edges = create_energy_bins(self.n_atoms, params.energy_min, params.energy_max, params.bins)
        self.comm = _AbiComm(self.dev, rank, world, unique_id) if comm == "abi" else _Comm(rank, world, torch_device)
        self.rng_local = np.random.default_rng([seed, rank])
        self.rng_shared = np.random.default_rng([seed, 0xEC])   # identical on all ranks
        # Philox key of this rank's Monte-Carlo kernels: the counters hold handle-local walker ids, so the rank goes into
        # the key (start states take it in the offset, _rand_offset)
        self.mc_seed = rank_seed(seed, rank)
        self.offset = 0
        q = np.repeat(np.arange(self.first_window, self.first_window + self.w_local), walkers)
        self.win_lo = self.window_indices[q, 0].astype(np.int32)
        self.win_hi = self.window_indices[q, 1].astype(np.int32)
        self.dev.wl_init(params.bins, self.edges, walkers)
        if self.span:
            self.dev.wl_set_span(world)
        self.dev.wl_set_windows(self.win_lo, self.win_hi, zero_hist=True)
        self.energies = np.zeros(self.n_local)
        self.hist_min, self.hist_mean = np.zeros(self.w_local), np.zeros(self.w_local)
        self.total_trials = 0
        self.stage_sweeps = []
        self.timing = {"enter": 0.0, "sweeps": 0.0, "collectives": 0.0, "exchange": 0.0, "stitch": 0.0,
                       "x_lng": 0.0, "x_plan": 0.0, "x_swap": 0.0}                                       # host seconds per part
        # load balancing (:1016-1019, 1079, 1108): trials each window spent unconverged, previous blend of weights
        self.wl_mc_steps = np.zeros(W)
        self.diffusion_prev = np.full(W, 1.0 / float(np.float32(W)))
        self.bin_width = energy_bin_width(self.n_atoms, params.energy_min, params.energy_max, params.bins)
        self.mean_energy = np.full((300, 2), 1.0 / (K_B_IN_RY * 10.0))          # :1112
        self.window_history = [self.window_indices.copy()]
        self.rand_calls = 0
        # rho(E): radial densities per energy bin (:574-592, save_rho_E :346-381); off when wc_range == 0
        self.wc_range, self.n_species = int(wc_range), n_species
        self.radial_record = np.zeros((self.n_local, params.bins), dtype=np.int64)
        self.radial_record_bool = np.zeros(params.bins, dtype=bool)
        self.rho_sum = np.zeros((params.bins, max(self.wc_range, 1), n_species, n_species))
        self.rho_saved, self.radial_min, self.rho_of_E = False, 0.0, None

    @property
    def lng(self):
        """ln g of every local walker [n_local][bins] (host copy; equal within a window after a `sweeps` call)."""
        return np.repeat(self.dev.wl_get(0), self.walkers, axis=0)

    @property
    def hist(self):
        return np.repeat(self.dev.wl_get(1), self.walkers, axis=0)

    def _set_windows(self, intervals):
        """mpi_arrays (:1351-1379): new overlapping index ranges for every walker of this rank."""
        self.intervals = np.array(intervals, dtype=np.int64)
        self.window_indices = create_overlap(self.intervals, self.p.bin_overlap)
        q = np.repeat(np.arange(self.first_window, self.first_window + self.w_local), self.walkers)
        self.win_lo = self.window_indices[q, 0].astype(np.int32)
        self.win_hi = self.window_indices[q, 1].astype(np.int32)
        self.dev.wl_set_windows(self.win_lo, self.win_hi, zero_hist=True)
        self.window_history.append(self.window_indices.copy())

    def _rand_offset(self):
        """Philox offset of the next batch of random start states: distinct per rank and per call (the replica index
        the device mixes in is local to the handle)."""
        self.rand_calls += 1
        return (0x5A << 56) | (self.rank << 32) | self.rand_calls

    # --- window entry (enter_energy_window, :643-741) ------------------------------------------------
    def enter_energy_windows(self, max_rounds=200, fresh=True):
        """fresh=False: start from the walkers' current configurations (after a window resize); walkers already
        inside their window leave the kernel at once."""
        import time
        t0 = time.perf_counter()
        try:
            return self._enter_energy_windows(max_rounds, fresh)
        finally:
            self.timing["enter"] += time.perf_counter() - t0

    def _enter_energy_windows(self, max_rounds, fresh):
        p = self.p
        lo = self.edges[self.win_lo - 1]
        hi = self.edges[self.win_hi]
        target = (lo + hi) / 2.0
        cond = np.abs(hi - lo) * 0.1
        sigma = float(np.float32(np.float32(0.0025) * np.float32(abs(p.energy_max - p.energy_min))) * np.float32(self.n_atoms)) \
            / (RY_TO_EV * 1000)                                # single-precision product, double division (:726-727)
        inv = 1.0 / (2.0 * sigma ** 2)
        pending = np.ones(self.n_local, dtype=bool)
        if fresh:                                              # start states generated on the device, all walkers at once
            self.dev.random_config(self.counts, 0, self.n_local, seed=self.seed, offset=self._rand_offset())
        for _ in range(max_rounds):
            e, ent = self.dev.wl_enter_window(target, lo + cond, hi - cond, inv, self.n_atoms * 250, self.mc_seed, self.offset)
            self.offset += 1
            pending = ent == 0
            if not pending.any():
                self.energies = e
                return
            off = self._rand_offset()
            for w in np.flatnonzero(pending):                 # re-randomise (:671-674)
                self.dev.random_config(self.counts, int(w), 1, seed=self.seed, offset=off)
        raise BrawlCudaError("walkers failed to enter their energy windows")

    # --- one outer iteration: sweeps + window average + replica exchange -------------------------------
    def _sweeps(self, wl_f):
        """sweeps (:539-626) for every local walker + the intra-window average (:628-631) + the flatness inputs, all on the
        device; the host receives the walkers' energies and (min, mean) of each window's histogram."""
        p = self.p
        n_trials = p.mc_sweeps * self.n_atoms
        self.energies, self.hist_min, self.hist_mean = self.dev.wl_iterate(wl_f, n_trials, seed=self.mc_seed, offset=self.offset,
                                                                           nbr_swap=p.nbr_swap)
        self.offset += 1
        self.total_trials += n_trials * self.n_local
        if self.wc_range and not self.rho_saved:
            self._sample_rho()

    def _sample_rho(self):
        """Radial densities of the walkers' configurations, accumulated in the energy bin they sit in (:574-592).  The
        reference samples inside the trial loop, at most once per n_atoms trials and in the bin of the proposed
        configuration; here the configuration at the end of each `sweeps` call is sampled (same estimator -- the mean
        of rho over configurations of a bin -- at a coarser cadence, ONE batched SRO launch per call).  Per walker and
        bin at most max(radial_samples / num_walkers, 1) samples, as in the reference."""
        cap = max(self.p.radial_samples // self.walkers_total, 1)
        want = []
        for w in range(self.n_local):
            jb = bin_index(self.energies[w], self.edges, self.p.bins)
            if 0 < jb < self.p.bins + 1 and not self.radial_record_bool[jb - 1] and self.radial_record[w, jb - 1] < cap:
                want.append((w, jb))
        if not want:
            return
        rho = self.dev.radial_densities_batch(self.wc_range, 0, self.n_local) if hasattr(self.dev, "radial_densities_batch") else None
        for w, jb in want:
            self.radial_record[w, jb - 1] += 1
            self.rho_sum[jb - 1] += rho[w] if rho is not None else self.dev.radial_densities(self.wc_range, w)

    def _save_rho_E(self):
        """save_rho_E (:346-381): a bin is complete once radial_samples samples exist over all ranks; when every bin is,
        rho_of_E[bin][shell][j][i] = the mean over its samples (what ncdf_radial_density_writer_across_energy stores)
        and sampling stops.  `rho_of_E_partial()` gives the same means over whatever has been sampled so far."""
        if not self.wc_range or self.rho_saved:
            return
        total = self.comm.all_gather(self.radial_record.sum(axis=0)).sum(axis=0)
        self.radial_record_bool |= total >= self.p.radial_samples
        self.radial_min = float(np.count_nonzero(self.radial_record_bool)) / float(self.p.bins)
        if self.radial_record_bool.all():
            self.rho_of_E = self.rho_of_E_partial()
            self.rho_saved = True

    def rho_of_E_partial(self):
        """(mean rho per bin [bins][wc_range][S][S] over all ranks, samples per bin [bins]); bins without samples are 0."""
        if self.rho_saved:
            return self.rho_of_E
        n = self.comm.all_gather(self.radial_record.sum(axis=0)).sum(axis=0)
        tot = self.comm.all_gather(self.rho_sum).sum(axis=0)
        return tot / np.maximum(n, 1)[:, None, None, None], n.astype(np.int64)

    def _window_lng_all(self):
        """Window-averaged ln g of all windows of all ranks: [W][bins] (dos_combine's gather, :1161-1192)."""
        if self.span:                                        # every rank holds every window, already averaged over all ranks
            return self.dev.wl_get(0)
        if self.world > 1 and getattr(self.comm, "abi", False):
            return self.dev.wl_allgather_lng(self.world).reshape(self.p.num_windows, self.p.bins)
        return self.comm.all_gather(self.dev.wl_get(0)).reshape(self.p.num_windows, self.p.bins)

    def _replica_exchange(self, e_all=None):
        if self.p.num_windows < 2 or self.p.performance not in (0, 2, 4):
            return 0
        import time
        if e_all is None:
            e_all = self.comm.all_gather(self.energies).reshape(-1)
        t0 = time.perf_counter()
        lng_all = self._window_lng_all()
        t1 = time.perf_counter()
        e_all = np.asarray(e_all)
        swaps = plan_replica_exchange(e_all[self.wm], lng_all, self.window_indices, self.walkers_total, self.edges, self.rng_shared)
        swaps = [(int(self.wm[a]), int(self.wm[b])) for a, b in swaps]          # window-major -> rank-major walker ids
        t2 = time.perf_counter()
        try:
            return self._do_swaps(swaps, e_all)
        finally:
            self.timing["x_lng"] += t1 - t0; self.timing["x_plan"] += t2 - t1; self.timing["x_swap"] += time.perf_counter() - t2

    def _do_swaps(self, swaps, e_all):
        """Carry out the accepted exchanges: same-GPU pairs in one launch, cross-GPU pairs in one NCCL group (a walker is
        matched at most once per call, so the pairs are disjoint)."""
        loc_a, loc_b, rem_r, rem_p = [], [], [], []
        for a, b in swaps:
            ra, rb = a // self.n_local, b // self.n_local
            la, lb = a % self.n_local, b % self.n_local
            if ra == rb == self.rank:
                loc_a.append(la); loc_b.append(lb)
                self.energies[la], self.energies[lb] = self.energies[lb], self.energies[la]
            elif self.rank in (ra, rb):
                mine, peer = (la, rb) if self.rank == ra else (lb, ra)
                rem_r.append(mine); rem_p.append(peer)
                self.energies[mine] = e_all[b if self.rank == ra else a]
        if loc_a:
            if hasattr(self.dev, "swap_replicas_batch"):
                self.dev.swap_replicas_batch(loc_a, loc_b)
            else:
                for la, lb in zip(loc_a, loc_b):
                    self.dev.swap_replicas(la, lb)
        if rem_r:
            if getattr(self.comm, "abi", False):
                self.dev.exchange_replicas(rem_r, rem_p)
            else:
                for mine, peer in zip(rem_r, rem_p):
                    self._exchange_remote(mine, peer)
        return len(swaps)

    def _exchange_remote(self, local_replica, peer_rank):
        """Swap one configuration with a walker on another GPU: device-to-device over NCCL."""
        if getattr(self.comm, "abi", False):
            self.dev.exchange_replica(local_replica, peer_rank)          # grouped ncclSend/ncclRecv on the handle's stream
            return
        torch, dist = self.comm.torch, self.comm.dist
        # the library's kernels run on the handle's own stream, torch's copies on torch's: order the two explicitly
        self.dev.synchronize()
        send = self.dev.lattice_tensor(torch)[local_replica].clone()
        recv = torch.empty_like(send)
        ops = [dist.P2POp(dist.isend, send, peer_rank), dist.P2POp(dist.irecv, recv, peer_rank)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        self.dev.lattice_tensor(torch)[local_replica].copy_(recv)
        if send.is_cuda:
            torch.cuda.current_stream().synchronize()

    def _flat_windows(self, min_hist=None):
        """flatness = minval(hist)/(sum(hist)/mpi_bins) > flatness per local window (:222-226), from the device's
        (min, mean) of the window-averaged histograms."""
        ok = []
        for q in range(self.w_local):
            mn, mean = float(self.hist_min[q]), float(self.hist_mean[q])
            good = (mn / mean if mean > 0 else 0.0) > self.p.flatness
            if min_hist is not None:
                good = good and mn > min_hist
            ok.append(bool(good))
        return ok

    def _stage(self, wl_f, min_hist=None, max_sweeps=100000, exchange_every=1, it=0, resize=False):
        converged = [False] * self.w_local
        n = 0
        per_sweep = float(self.p.mc_sweeps * self.n_atoms * self.walkers)        # every walker adds its trials (:217, :775)
        import time
        while True:
            n += 1
            t0 = time.perf_counter()
            self._sweeps(wl_f)
            t1 = time.perf_counter()
            for q in range(self.w_local):
                if not converged[q]:
                    self.wl_mc_steps[self.first_window + q] += per_sweep
            flat = self._flat_windows(min_hist)
            converged = [c or f for c, f in zip(converged, flat)]
            # one collective per iteration: walker energies (replica_exchange, :1435) + the converged count (:230)
            both = self.comm.all_gather(np.concatenate([self.energies, [float(sum(converged))]]))
            t2 = time.perf_counter()
            if n % exchange_every == 0:
                self._replica_exchange(both[:, :-1].reshape(-1))
            t3 = time.perf_counter()
            self.timing["sweeps"] += t1 - t0; self.timing["collectives"] += t2 - t1; self.timing["exchange"] += t3 - t2
            # span: every rank sees the same (all-reduced) histograms and reports all windows
            if (both[0, -1] if self.span else both[:, -1].sum()) == self.p.num_windows or n >= max_sweeps:
                break
        self.stage_sweeps.append(n)
        t4 = time.perf_counter()
        self._save_rho_E()
        combined = dos_combine(self._window_lng_all(), self.window_indices)      # dos_average + dos_combine
        self.dev.wl_set_lng(combined)
        self.dev.wl_zero_hist()
        steps = self.comm.all_gather(self.wl_mc_steps).sum(axis=0)               # MPI_ALLREDUCE(wl_mc_steps) (:244, :814)
        self.last_mc_steps = steps.copy()
        if resize and self.p.num_windows > 1:
            # every rank evaluates the same arithmetic on the same inputs (the reference computes on rank 0 and broadcasts)
            iv, self.diffusion_prev = window_optimise(it, self.intervals, steps, self.diffusion_prev, self.p.bins)
            self._set_windows(iv)
        self.wl_mc_steps[...] = 0.0
        self.mean_energy = compute_mean_energy(combined, self.edges, self.p.bins, self.bin_width)
        self.timing["stitch"] += time.perf_counter() - t4
        if resize and self.p.num_windows > 1:
            self.enter_energy_windows(fresh=False)                               # in place of load_window_config
        return combined

    def save_wl_data(self, directory, lng, hist=None):
        """save_wl_data (:409-417), rank 0 only: data/wl_dos_bins.nc (bin edges), data/wl_dos.nc (ln g), data/wl_hist.nc."""
        if self.rank != 0:
            return
        import os
        d = os.path.join(directory, "data")
        os.makedirs(d, exist_ok=True)
        ncdf_writer_1d(os.path.join(d, "wl_dos_bins.nc"), self.edges)
        ncdf_writer_1d(os.path.join(d, "wl_dos.nc"), lng)
        ncdf_writer_1d(os.path.join(d, "wl_hist.nc"), np.zeros(self.p.bins) if hist is None else hist)

    def save_rho_of_E(self, directory, shells, setup):
        """The file save_rho_E writes once every bin is complete (:374-391): asro/rho_of_E.nc with rho(i,j,r,bin), the
        shell radii (lattice_shells) and bin_energy(i) = E_min + (i - 0.5) * bin_width (:985).  `setup`: the attribute
        fields of netcdf3._setup_atts.  Writes the partial means if sampling has not completed; rank 0 only."""
        rho, _ = self.rho_of_E_partial()
        if self.rank != 0:
            return
        import os
        d = os.path.join(directory, "asro")
        os.makedirs(d, exist_ok=True)
        e_min = self.p.energy_min * (self.n_atoms / (RY_TO_EV * 1000))
        bin_energy = np.array([e_min + (i - 0.5) * self.bin_width for i in range(1, self.p.bins + 1)])
        ncdf_radial_density_writer_across_energy(os.path.join(d, "rho_of_E.nc"), rho, shells, bin_energy, setup)

    def run(self, max_sweeps_per_stage=100000, callback=None):
        """pre_sampling (:757-838) then the f-halving loop (:198-292).  Returns ln g(E) [bins]."""
        p = self.p
        self.enter_energy_windows()
        wl_f = p.wl_f
        combined = self._stage(wl_f, min_hist=1000.0 / float(np.float32(self.walkers_total)), max_sweeps=max_sweeps_per_stage,
                               exchange_every=10, it=0, resize=p.performance in (0, 1, 2, 3))
        it = 1
        while wl_f > p.tolerance:
            combined = self._stage(wl_f, max_sweeps=max_sweeps_per_stage, it=it, resize=p.performance in (0, 1))
            it += 1
            wl_f = wl_f * 0.5
            if callback:
                callback(wl_f, combined)
        return combined
