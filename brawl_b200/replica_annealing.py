"""Replica-batched simulated annealing on the C ABI -- BASELINE config 5 (1024 independent replicas per GPU for
SRO-vs-T sweeps) and SURVEY 8(e) "replica-batched Metropolis".

The reference parallelises Metropolis by running one independent chain per MPI rank
(metropolis_simulated_annealing, src/metropolis.F90:193-447) and averaging the per-rank results at the end
(comms_reduce_metropolis_results, src/comms.F90:122-160: MPI_SUM, then / p).  Here a chain is a replica: every GPU
holds `n_replicas` lattices in one handle, all replicas of a GPU advance in the same kernel launches
(brawl_cuda_metropolis_run with a per-replica beta), energies come from one batched total_energy launch and the SRO
pair counts from the radial-counts kernel.  Ranks never talk during the run; the only collective is the end-of-run
sum of T_steps x (3 + S^2 wc_range) float64 over ranks (torch.distributed all_gather: NCCL on GPUs, gloo in the CPU
test), exactly the reduction the reference does.

Loop structure, sampling cadence and the formulas for <E>, C and the acceptance rate follow the reference line by
line (cited below); what differs is the proposal stream (Philox, production kernels) and therefore the agreement is
statistical (tests/test_gpu_parity.py::test_production_statistics_match_oracle pins the sampler itself).
"""
import math

import numpy as np

from .engine import Device, K_B_IN_RY, BrawlCudaError, rank_seed


def reduce_results(results, comm_all_gather, world):
    """comms_reduce_metropolis_results (src/comms.F90:122-160): sum over ranks, divide by the number of chains.
    `results`: dict of per-rank arrays already summed over the rank's replicas + "n_chains"."""
    out = {}
    n = float(comm_all_gather(np.array([float(results["n_chains"])])).sum())
    for k, v in results.items():
        if k == "n_chains":
            continue
        out[k] = comm_all_gather(np.asarray(v, dtype=np.float64)).sum(axis=0) / n
    out["n_chains"] = int(n)
    return out


class ReplicaAnnealing:
    def __init__(self, lattice, n_1, n_2, n_3, n_species, n_shells, V_ex, counts, n_replicas, T, T_steps, delta_T,
                 n_mc_steps, n_sample_steps, n_burn_in_steps=0, burn_in_start=False, burn_in=False,
                 n_sample_steps_asro=None, wc_range=2, nbr_swap=False, device=0, rank=0, world=1, seed=0x42726157,
                 torch_device=None, device_cls=None):
        if n_mc_steps < 1 or n_sample_steps < 1 or n_mc_steps < n_sample_steps:
            raise BrawlCudaError("need n_mc_steps >= n_sample_steps >= 1")
        # device_cls: test hook (the CPU tests stand the oracle in for the CUDA handle); never set by the product
        self.dev = (device_cls or Device)(lattice, n_1, n_2, n_3, n_species, n_shells, V_ex, device=device, n_replicas=n_replicas)
        self.R, self.S, self.counts = n_replicas, n_species, list(counts)
        self.T, self.T_steps, self.delta_T = float(T), int(T_steps), float(delta_T)
        self.n_mc_steps, self.n_sample_steps = int(n_mc_steps), int(n_sample_steps)
        self.n_burn_in_steps, self.burn_in_start, self.burn_in = int(n_burn_in_steps), bool(burn_in_start), bool(burn_in)
        self.n_sample_steps_asro = int(n_sample_steps_asro or n_sample_steps)          # io.f90:654-662
        self.wc_range, self.nbr_swap = int(wc_range), bool(nbr_swap)
        self.rank, self.world, self.seed = rank, world, seed
        # the Philox counters of the Monte-Carlo kernels hold (trial, tag, handle-local replica / box, phase): the rank
        # goes into the key, or replica j of every rank would draw the same proposals and uniforms
        self.mc_seed = rank_seed(seed, rank)
        from .wang_landau import _Comm
        self.comm = _Comm(rank, world, torch_device)
        self.n_atoms = self.dev.n_atoms

    def _radial_densities(self):
        """r_densities(i,j,l) per replica: [R][wc_range][S][S] (analytics.f90:293-404)."""
        return self.dev.radial_densities_batch(self.wc_range, 0, self.R)      # one launch for all replicas

    def run(self, initial_configs=None):
        """Returns (per_replica, averaged): per_replica = dict of arrays [R][T_steps]...; averaged = the reference's
        av_* outputs over all chains of all ranks."""
        dev, R, N = self.dev, self.R, self.n_atoms
        if initial_configs is None:
            dev.random_config(self.counts, 0, R, seed=self.seed, offset=(0x5B << 56) | (self.rank << 32))
        else:
            dev.set_config(initial_configs, 0, R)
        n_save_energy = math.floor(np.float32(self.n_mc_steps) / np.float32(self.n_sample_steps))           # :160-161
        n_save_asro = math.floor(np.float32(self.n_mc_steps) / np.float32(self.n_sample_steps_asro))        # :162-163
        temperature = np.zeros(self.T_steps)
        energies_of_T = np.zeros((R, self.T_steps))
        C_of_T = np.zeros((R, self.T_steps))
        acceptance_of_T = np.zeros((R, self.T_steps))
        rho_of_T = np.zeros((R, self.T_steps, self.wc_range, self.S, self.S))
        self.attempted = 0
        for j in range(1, self.T_steps + 1):
            temp = self.T + float(j - 1) * self.delta_T                                                   # :204
            sim_temp = temp * K_B_IN_RY
            beta = 1.0 / sim_temp if sim_temp != 0.0 else math.inf
            temperature[j - 1] = temp
            if (j == 1 and self.burn_in_start) or (j > 1 and self.burn_in):                               # :214-238
                att, _, _ = dev.metropolis_run(beta, self.n_burn_in_steps, seed=self.mc_seed, nbr_swap=self.nbr_swap)
                self.attempted += int(att.sum())
            n_sweeps = self.n_mc_steps // self.n_sample_steps                                             # :343-344
            n_sweep_steps = self.n_mc_steps // n_sweeps
            step_E, step_Esq = np.zeros(R), np.zeros(R)
            accepted, attempted = np.zeros(R), np.zeros(R)
            r_dens = np.zeros((R, self.wc_range, self.S, self.S))
            for i in range(1, n_sweeps + 1):
                att, acc, _ = dev.metropolis_run(beta, n_sweep_steps, seed=self.mc_seed, nbr_swap=self.nbr_swap)   # :350-354
                accepted += acc
                attempted += att
                step_n = i * n_sweep_steps
                e = dev.total_energy(0, R, exact_order=False)                                             # :361
                step_E += e
                step_Esq += e * e
                if self.wc_range and step_n % self.n_sample_steps_asro == 0:                              # :378-383
                    r_dens += self._radial_densities()
            self.attempted += int(attempted.sum())
            # the production kernels attempt whole sweeps of their decomposition (>= the trials asked for), so the rate is
            # accepted / attempted rather than / n_mc_steps (:412)
            acceptance_of_T[:, j - 1] = accepted / np.maximum(attempted, 1.0)
            energies_of_T[:, j - 1] = step_E / n_save_energy / N                                          # :416
            if temp > 0.0:
                C = (step_Esq / n_save_energy - (step_E / n_save_energy) ** 2) / (sim_temp * temp) / N    # :419-424
                C_of_T[:, j - 1] = np.maximum(C, 0.0)
            if self.wc_range and n_save_asro > 0:
                rho_of_T[:, j - 1] = r_dens / n_save_asro                                                 # :441
        per = dict(temperature=temperature, energies_of_T=energies_of_T, C_of_T=C_of_T, acceptance_of_T=acceptance_of_T,
                   rho_of_T=rho_of_T)
        local = dict(n_chains=R, energies_of_T=energies_of_T.sum(axis=0), C_of_T=C_of_T.sum(axis=0),
                     acceptance_of_T=acceptance_of_T.sum(axis=0), rho_of_T=rho_of_T.sum(axis=0))
        av = reduce_results(local, self.comm.all_gather, self.world)
        av["temperature"] = temperature
        return per, av


def save_av_radial_density(directory, av, shells, setup):
    """asro/av_radial_density.nc as metropolis_simulated_annealing writes it after the rank average (metropolis.F90:529;
    ncdf_radial_density_writer, netcdf_io.f90:150-244): rho(i,j,r,T), shell radii, temperatures, <E>(T) per atom."""
    import os
    from .netcdf3 import ncdf_radial_density_writer
    d = os.path.join(directory, "asro")
    os.makedirs(d, exist_ok=True)
    ncdf_radial_density_writer(os.path.join(d, "av_radial_density.nc"), av["rho_of_T"], shells, av["temperature"],
                               av["energies_of_T"], setup)


def save_av_energy_diagnostics(directory, av):
    """energies/av_energy_diagnostics.dat (diagnostics_writer, src/metropolis_output.f90:113-135)."""
    import os
    from .text_io import diagnostics_writer
    d = os.path.join(directory, "energies")
    os.makedirs(d, exist_ok=True)
    diagnostics_writer(os.path.join(d, "av_energy_diagnostics.dat"), av["temperature"], av["energies_of_T"], av["C_of_T"],
                       av["acceptance_of_T"])


def warren_cowley(rho, concentrations, coordination):
    """alpha^{ij}_n = 1 - rho^{ij}_n / (Z_n c_j) (examples/01_metropolis_FeNi/02_simulated_annealing/01_plot_results.py:35-36);
    rho[..., shell, j, i], coordination[shell] (use 1 for the r = 0 shell)."""
    c = np.asarray(concentrations, dtype=np.float64)
    z = np.asarray(coordination, dtype=np.float64)
    return 1.0 - rho / (z[:, None, None] * c[None, :, None])
