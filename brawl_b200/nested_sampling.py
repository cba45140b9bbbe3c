"""Nested sampling on top of the C ABI -- host-side mirror of nested_sampling_main
(src/nested_sampling.f90:45-206) with the constrained random walk (:157-192), the walker clone
(:151) and the initial energies (:95) on the GPU.

The reference algorithm is serial inside one run (cull the highest-energy walker, clone a random
survivor, walk the clone under the energy ceiling).  Batching therefore happens ACROSS runs: `n_runs`
independent nested-sampling runs (each with its own K walkers) advance in lock-step, and every
iteration walks n_runs clones concurrently with one kernel launch (one warp per walking clone).
n_runs = 1 with `mt` given replays the reference's MT19937 stream bit-for-bit (golden case 03).
"""
import math

import numpy as np

from .engine import Device, BrawlCudaError


class NSParams:
    """ns_params (src/derived_types.f90:242-255; read_ns_file src/io.f90:740-817)."""

    def __init__(self, n_walkers=100, n_steps=500, n_iter=1000, traj_freq=100, outfile_ener="ns.energies",
                 outfile_traj="ns.traj.xyz"):
        self.n_walkers, self.n_steps, self.n_iter, self.traj_freq = int(n_walkers), int(n_steps), int(n_iter), int(traj_freq)
        self.outfile_ener, self.outfile_traj = outfile_ener, outfile_traj

    @classmethod
    def from_file(cls, path):
        kw = {}
        for line in open(path):
            if "=" not in line or line.lstrip().startswith("#"):
                continue
            k, v = line.split("=", 1)
            k, v = k.rstrip(), v.split("#")[0].strip().strip("'\"")
            if k in ("n_walkers", "n_steps", "n_iter", "traj_freq"):
                kw[k] = int(v)
            elif k in ("outfile_ener", "outfile_traj"):
                kw[k] = v
        return cls(**kw)


class NestedSampling:
    def __init__(self, lattice, n_1, n_2, n_3, n_species, n_shells, V_ex, counts, params, n_runs=1, device=0,
                 seed=0x42726157):
        self.p, self.R, self.K, self.seed = params, int(n_runs), params.n_walkers, seed
        self.dev = Device(lattice, n_1, n_2, n_3, n_species, n_shells, V_ex, device=device, n_replicas=self.R * self.K)
        self.lattice, self.n, self.counts, self.S = lattice, (n_1, n_2, n_3), counts, n_species
        self.rng = np.random.default_rng([seed, 0x25])
        # n_at as the reference computes it (nested_sampling.f90:84): n_1*n_2*n_3*n_basis*n_species (sic)
        self.n_at = n_1 * n_2 * n_3 * 1 * n_species
        self.energies = None

    def initialise(self, configs=None):
        """Random initial walkers (:78-97) unless `configs[R*K]` is given; E = full_energy + u*1e-8
        (the 1e-8 is a default-real literal, :95)."""
        n = self.R * self.K
        if configs is None:                                    # all R*K start states in one device call
            self.dev.random_config(self.counts, 0, n, seed=self.seed, offset=0x5C << 56)
        else:
            self.dev.set_config(configs, 0, n)
        e = self.dev.total_energy(0, n, exact_order=True)
        self.energies = (e + self.rng.random(n) * float(np.float32(1e-8))).reshape(self.R, self.K)

    def run(self, n_iter=None, callback=None):
        """Returns culled[R][n_iter]: the energy ceilings written to the .energies file (:121)."""
        p = self.p
        n_iter = p.n_iter if n_iter is None else n_iter
        if self.energies is None:
            self.initialise()
        R, K = self.R, self.K
        culled = np.zeros((R, n_iter))
        extra = np.zeros(R, dtype=np.int64)
        n_acc = np.zeros(R, dtype=np.int64)
        base = np.arange(R) * K
        for it in range(1, n_iter + 1):
            i_max = np.argmax(self.energies, axis=1)                     # maxloc: first maximum
            lim = self.energies[np.arange(R), i_max]
            culled[:, it - 1] = lim
            if it % max(int(K / 2.0), 1) == 0:                           # :129-144 (n_walkers = 1: the reference divides by zero)
                grow = (n_acc < self.n_at * np.float32(0.05)) & (extra < p.n_steps * 100)
                extra[grow] += p.n_steps
            irnd = np.maximum(np.ceil(self.rng.random(R) * K).astype(np.int64), 1)     # :149-150
            self.dev.copy_replicas_batch(base + irnd - 1, base + i_max)                   # :151, one launch for all runs
            self.energies[np.arange(R), i_max] = self.energies[np.arange(R), irnd - 1]
            # all runs share the step count of the run that needs most (extra steps never hurt:
            # the walk is a valid constrained random walk of any length)
            steps = int(p.n_steps + extra.max())
            e_new, acc = self.dev.ns_walk(base + i_max, self.energies[np.arange(R), i_max], lim, steps,
                                          seed=self.seed, offset=it)
            self.energies[np.arange(R), i_max] = e_new
            n_acc = acc
            if callback:
                callback(it, lim)
        return culled

    def write_energies(self, path, culled_run):
        """The .energies file of the reference (header :87, one '(i_iter, ener_limit)' line per
        iteration :121; list-directed output: 17 significant digits)."""
        with open(path, "w") as fh:
            fh.write("%12d%12d%12d False%12d\n" % (self.K, 1, 0, self.n_at))
            for i, e in enumerate(culled_run):
                fh.write("%12d  %s     \n" % (i + 1, repr(float(e))))
