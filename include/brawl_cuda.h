/*
 * brawl_cuda.h -- C ABI of libbrawl_cuda.so: BraWl's atom-swap Monte-Carlo hot path on B200.
 *
 * This is the drop-in boundary.  Every entry point is `extern "C"`, takes plain pointers and
 * sizes, returns an int status (0 = OK; non-zero = error, text via brawl_cuda_last_error())
 * and never throws, aborts or falls back to a CPU path: if no CUDA device is usable the
 * call fails.  It extends the reference's only FFI seam, src/c_functions.f90:21-37 (the
 * ISO_C_BINDING interface block for genrand / f90_init_genrand); INTEGRATION.md shows the
 * Fortran `interface` blocks and the shims that rebind the run_params operator table
 * (src/derived_types.f90:90-98) to these functions.
 *
 * Conventions shared by all calls
 *  - lattice_id: 0 = simple_cubic, 1 = bcc, 2 = fcc   (setup%lattice, src/initialise.F90:159-240)
 *  - A configuration is the reference's own array: int8 config(1, 2*n_1, 2*n_2, 2*n_3),
 *    column-major, i.e. a contiguous byte grid[z][y][x] with x fastest, species 1..n_species,
 *    0 = "no lattice site here" (src/shared_data.f90:30, src/initialise.F90:295).  Fortran passes
 *    c_loc(config).  Inside the library the lattice lives in HBM in a compact sites-only layout.
 *  - V_ex is the reference's array V_ex(n_species, n_species, n_shells) exactly as read from the
 *    *.vij file (src/io.f90:401): V_ex[(shell*S + nbr)*S + centre], f64, Rydberg.
 *  - Site coordinates / flat indices are 0-based on the doubled grid: idx = (z*2n_2 + y)*2n_1 + x
 *    (Fortran site (1,i,j,k) -> x=i-1, y=j-1, z=k-1).
 *  - A handle owns `n_replicas` independent lattices of the same shape (one per MPI rank /
 *    Wang-Landau walker / nested-sampling walker in the reference); replica r of batched
 *    host arrays starts at r * (8*n_1*n_2*n_3) bytes.
 *  - All work is enqueued on the handle's CUDA stream; calls that return data to host memory
 *    synchronise that stream before returning.
 */
#ifndef BRAWL_CUDA_H
#define BRAWL_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct brawl_cuda_ctx brawl_cuda_t;

#define BRAWL_LATTICE_SC 0
#define BRAWL_LATTICE_BCC 1
#define BRAWL_LATTICE_FCC 2

/* ---- library / error state --------------------------------------------------------------- */
const char *brawl_cuda_last_error(void);        /* thread-local text of the last failure      */
int brawl_cuda_version(void);                   /* ABI version, currently 1                    */
int brawl_cuda_device_count(int *count);        /* fails (non-zero) if the CUDA runtime cannot start */

/* Self-test hook: one Philox4x32-10 block computed ON THE DEVICE (counter[4], key[2] -> out[4]);
 * lets the tests pin the production RNG to the published known-answer vectors. */
int brawl_cuda_philox4x32(int device, const uint32_t *counter4, const uint32_t *key2, uint32_t *out4);

/* ---- handle -------------------------------------------------------------------------------
 * Replaces: initialise_function_pointers (src/initialise.F90:153-257: binds nbr_energy by
 * lattice x interaction_range), initialise_interaction/read_exchange (V_ex upload,
 * src/initialise.F90:268-275, src/io.f90:389-415) and initialise_local_arrays (config
 * allocation, src/initialise.F90:286-300).  Unsupported lattice/shell combinations fail the
 * same way the reference does ("Unsupported number of shells"). */
int brawl_cuda_create(int lattice_id, int n_1, int n_2, int n_3, int n_species, int n_shells,
                      const double *V_ex, int device, int n_replicas, brawl_cuda_t **handle);
int brawl_cuda_destroy(brawl_cuda_t *h);
/* Launch on an existing stream (a cudaStream_t cast to void*; NULL = the library's own) */
int brawl_cuda_set_stream(brawl_cuda_t *h, void *cuda_stream);
int brawl_cuda_synchronize(brawl_cuda_t *h);
int brawl_cuda_info(brawl_cuda_t *h, int64_t *n_atoms, int64_t *grid_bytes, int *z_total, int *n_replicas);

/* ---- configuration transfer ----------------------------------------------------------------
 * Host <-> HBM copy of the reference-layout grid(s) (what `config` is to every reference
 * routine).  set validates occupancy: a byte that is 0 on a lattice site, non-zero off-site or
 * > n_species is an error. */
int brawl_cuda_set_config(brawl_cuda_t *h, int first_replica, int n, const int8_t *grids);
int brawl_cuda_get_config(brawl_cuda_t *h, int first_replica, int n, int8_t *grids);
/* The same configurations as compact host buffers: sites[n][n_atoms] bytes, species 0..S-1 (NOT 1..S) in the device's
 * site order (z slowest, then compact y, compact x: site (x, y, z) of the doubled grid -> bcc ((z*n2 + y/2)*n1 + x/2),
 * fcc ((z*2*n2 + y)*n1 + x/2)); 1/4 (bcc) or 1/2 (fcc) of the bytes of the reference's config grid
 * (src/shared_data.f90:30).  For drivers that keep the configuration between annealing segments / in checkpoints and
 * share PCIe or host-memory bandwidth between many GPUs.  set_lattice fails if a byte is >= n_species. */
int brawl_cuda_set_lattice(brawl_cuda_t *h, int first_replica, int n, const uint8_t *sites);
int brawl_cuda_get_lattice(brawl_cuda_t *h, int first_replica, int n, uint8_t *sites);
/* replica dst := replica src on the device (nested_sampling.f90:151 walker cloning;
 * wang-landau.F90:1475-1495 same-GPU replica exchange) */
int brawl_cuda_copy_replica(brawl_cuda_t *h, int src, int dst);
/* Atomic long-range order.  store_state == analytics.f90:43-64 (called from metropolis.F90:397 every
 * n_sample_steps_alro trials): for replicas [first_replica, first_replica+n) add 1 to the occupancy count of the
 * species sitting on every site.  The counts live on the device (uint32, [replica][species][site], allocated on the
 * first call).  get_order writes them in the reference's layout order(species, 1, x, y, z) -- species fastest, then the
 * (2n1, 2n2, 2n3) grid, off-site cells 0 -- as float64 counts (the caller divides by the number of samples,
 * metropolis.F90:446); reset != 0 zeroes that replica's counts afterwards (one temperature done). */
int brawl_cuda_store_state(brawl_cuda_t *h, int first_replica, int n);
int brawl_cuda_get_order(brawl_cuda_t *h, int replica, double *order, int reset);
/* Start states generated on the device: replicas [first_replica, first_replica+n) each become an independent,
 * uniformly random arrangement of the species multiset species_count[n_species] (which must sum to the number of
 * lattice sites).  Stands in for initial_setup (src/initialise.F90:434-617) called once per replica and for the
 * re-randomisation inside enter_energy_window (src/wang-landau.F90:671-674) when thousands of replicas / walkers are
 * started at once; the reference's own MT-stream fill stays a host routine (the drivers' initial_setup), this entry
 * draws from Philox (seed, offset, replica, site).  Deterministic for given (seed, offset, first_replica + i). */
int brawl_cuda_random_config(brawl_cuda_t *h, int first_replica, int n, const int64_t *species_count, uint64_t seed,
                             uint64_t offset);

/* ---- Hamiltonian ----------------------------------------------------------------------------
 * total_energy == setup%full_energy (src/bw_hamiltonian.f90:58-81).  exact_order != 0 adds the
 * per-site energies in the reference's z/y/x order => bit-identical to the reference;
 * exact_order == 0 uses a parallel tree (same per-site values, different last-bit rounding).
 * energies[n] for replicas first_replica .. first_replica+n-1. */
int brawl_cuda_total_energy(brawl_cuda_t *h, int first_replica, int n, int exact_order, double *energies);
/* nbr_energy == setup%nbr_energy(config, 1, x+1, y+1, z+1) (src/bw_hamiltonian.f90:898-1262,
 * 1712-1910, 2059-2106) for every cell of one replica; 0.0 on empty cells.  out[grid]. */
int brawl_cuda_site_energies(brawl_cuda_t *h, int replica, double *out);
/* The operator itself, one site: setup%nbr_energy(config, 1, x+1, y+1, z+1) with the reference's summation order
 * (bit-identical).  species = 0: the occupant is the centre; species = s > 0: species s is (the Fortran functions take the
 * centre species from config after pair_swap, src/metropolis.F90:783-792 -- this argument spares the swap). */
int brawl_cuda_nbr_energy(brawl_cuda_t *h, int replica, int x, int y, int z, int species, double *energy);
/* Per-swap dE exactly as monte_carlo_step_* forms it: pair_energy(after) - pair_energy(before)
 * (src/metropolis.F90:783-792, src/bw_hamiltonian.f90:99-114); configuration is not modified. */
int brawl_cuda_pair_dE(brawl_cuda_t *h, int replica, int64_t n_pairs, const int32_t *idx1,
                       const int32_t *idx2, double *dE);

/* ---- Metropolis -----------------------------------------------------------------------------
 * Deterministic replay: n_trials calls of setup%mc_step (monte_carlo_step_lattice / _nbr,
 * src/metropolis.F90:751-891) on one replica, consuming the reference's MT19937 stream
 * (src/mt19937ar.c; mt_state = mt[624] followed by mti, updated in place) in the reference's
 * order (z, x, y per site; +1 draw only if species differ and dE >= 0).  Replaces the k-loop at
 * src/metropolis.F90:350-354.  Bit-exact trajectory. */
int brawl_cuda_metropolis_replay(brawl_cuda_t *h, int replica, double beta, int64_t n_trials,
                                 int nbr_swap, uint32_t *mt_state625, int64_t *n_accept);
/* Same, with the sampling loop fused: every n_sample_steps trials the exact-order total energy
 * is recorded on the device (src/metropolis.F90:348-367); energies[n_trials/n_sample_steps]. */
int brawl_cuda_metropolis_replay_sampled(brawl_cuda_t *h, int replica, double beta, int64_t n_trials,
                                         int64_t n_sample_steps, int nbr_swap, uint32_t *mt_state625,
                                         int64_t *n_accept, double *energies);

/* Production: every replica performs >= n_trials attempted swaps at its own beta[r] with
 * counter-based Philox4x32-10 streams (key = seed, counter = (site-pair slot, step, box, phase)).
 * Large lattices run as shared-memory boxes with sublattice-parallel swaps, small ones as one
 * sequential chain per replica.  Outputs (any may be NULL), per replica:
 * attempted, accepted (same-species proposals count as accepted, src/metropolis.F90:774-777) and the
 * sum of accepted dE.  `offset` must advance between calls that share a seed: pass the returned
 * *next_offset. */
int brawl_cuda_metropolis_run(brawl_cuda_t *h, const double *beta, int64_t n_trials, int nbr_swap,
                              uint64_t seed, uint64_t offset, uint64_t *next_offset,
                              int64_t *n_attempt, int64_t *n_accept, double *dE_sum);
/* Enqueue-only form (no host sync, counters accumulate on the device until
 * brawl_cuda_metropolis_counters is called) -- what the benchmark times with CUDA events. */
int brawl_cuda_metropolis_enqueue(brawl_cuda_t *h, const double *beta, int64_t n_trials, int nbr_swap,
                                  uint64_t seed, uint64_t offset, uint64_t *next_offset,
                                  int64_t *n_attempt_planned, int *n_kernel_launches);
int brawl_cuda_metropolis_counters(brawl_cuda_t *h, int reset, int64_t *n_attempt, int64_t *n_accept,
                                   double *dE_sum);
/* number of kernel launches of the last metropolis_run / metropolis_enqueue on this handle */
int brawl_cuda_metropolis_last_launches(brawl_cuda_t *h, int *n_launches);
/* Tuning of the box decomposition (0 = automatic): box extents in doubled-grid units, trial
 * steps per phase.  Test hooks: steps_per_phase = -(s+1) selects s steps AND forces the generic
 * runtime-geometry kernel instead of a specialised instantiation; adding 100000 to the (positive)
 * step count restricts the planner to cubic periods (P,P,P). */
int brawl_cuda_metropolis_tune(brawl_cuda_t *h, int box_x, int box_y, int box_z, int steps_per_phase);
/* How the specialised box kernels form dE.  0: the reference's f64 association for every trial.
 * 1: integer neighbour counts (byte lattice in shared memory) give dE first; any trial whose dE or
 * acceptance test lies within a guard band of a decision boundary is recomputed with the reference's
 * association and decided by it, so accept/reject decisions -- and therefore trajectories -- are identical
 * to mode 0 on the same decomposition (needs <= 5 species; otherwise mode 0 is used).  2 (default): the same
 * screening on a word lattice (one 32-bit word per site, fixed-point dp4a dE, ex2.approx acceptance pre-test)
 * with the dense non-interacting-set decomposition (bcc, 4 shells, 64x64x32 boxes); where no word kernel is
 * instantiated it behaves like mode 1.  Where a word kernel exists, mode 0 runs its EXACT instantiation (same
 * decomposition, reference association for every trial), so modes 0 and 2 give identical trajectories; mode 1
 * keeps the byte-lattice decomposition.  The returned sum of accepted dE is exact to f64 rounding in modes 0/1
 * and to the fixed-point unit in mode 2 (epoch kernels: 22-bit table entries, < 2e-10 Ry per accepted swap; the
 * decisions themselves are exact in every mode). */
int brawl_cuda_metropolis_set_mode(brawl_cuda_t *h, int dE_mode);
/* byte_layout_only: 0 (default) automatic: where instantiated (bcc, 4 shells, <= 5 species) the epoch kernel with the site
 * energies cached over epochs of 4 steps (epoch_metropolis.cuh); 4 / 5: the same kernel with epochs of 8 / 2 steps (more
 * attempts per second, fewer of them effective / the reverse: DESIGN.md 4.3); 3: the one-gather-per-step word kernel with two
 * warp groups (the round-1 default), 2: without the split and the shared z margins; 1: never use the word-lattice kernels and
 * their dense decomposition (A/B comparisons, tests). */
int brawl_cuda_metropolis_set_layout(brawl_cuda_t *h, int byte_layout_only);
/* Describe the decomposition chosen: out10 = { kind + 16*n_orientations, period Px*10000+Py*100+Pz of the
 * first orientation, margin, box_x, box_y, box_z, max trials per step, boxes per replica, |D| (allowed
 * displacement classes of the first orientation), steps per phase }.  kind: 0 chain kernel, 1 generic box
 * kernel, 2 specialised (compile-time geometry) box kernel, 3 specialised + screened dE, 4 word-lattice
 * screened kernel (dense decomposition), 5 word-lattice kernel deciding every trial with the reference association.
 * out10[0] bit 12: the word kernel runs two independent warp groups per CTA; box_z is then the z pitch of the box
 * layers (box depth = pitch + margin: consecutive layers share their frozen margin planes). */
int brawl_cuda_metropolis_plan(brawl_cuda_t *h, int nbr_swap, int *out10);

/* ---- short-range order ----------------------------------------------------------------------
 * Integer pair counts behind radial_densities (src/analytics.f90:293-404):
 * cnt[(l*S + j)*S + i] = number of species-(j+1) atoms on coordination shell l of species-(i+1)
 * atoms (l = 0: the site itself), species_count[S].  rho = cnt / species_count[i] in f64 is
 * exactly the reference's r_densities(i,j,l).  wc_range-1 <= n tabulated shells. */
int brawl_cuda_radial_counts(brawl_cuda_t *h, int replica, int wc_range, int64_t *cnt, int64_t *species_count);
/* the same for replicas [first_replica, first_replica+n) in one launch: cnt[n][wc_range][S][S], species_count[n][S]
 * (SRO of every chain of a replica batch per sampling point, metropolis.F90:378-383) */
int brawl_cuda_radial_counts_batch(brawl_cuda_t *h, int first_replica, int n, int wc_range, int64_t *cnt,
                                   int64_t *species_count);

/* ---- Wang-Landau ----------------------------------------------------------------------------
 * sweeps() for the walkers held by this handle (src/wang-landau.F90:539-626): walker w owns
 * replica w, lng[w*bins .. ], hist[w*hist_stride ..], window win_lo[w]..win_hi[w] (1-based bin
 * indices mpi_start_idx/mpi_end_idx).  The intra-window average (:628-631) is done by
 * brawl_cuda_wl_window_average for walkers on this GPU and by the caller's NCCL allreduce
 * across GPUs.  Replay form consumes an MT19937 stream for ONE walker (bit-exact). */
int brawl_cuda_wl_sweeps_replay(brawl_cuda_t *h, int replica, double *lng, double *hist, const double *bin_edges,
                                int bins, int win_lo, int win_hi, double wl_f, int64_t n_trials, int nbr_swap,
                                uint32_t *mt_state625, int64_t *n_accept, double *e_final);
int brawl_cuda_wl_sweeps(brawl_cuda_t *h, int n_walkers, double *lng_dev_or_host, double *hist_dev_or_host,
                         int data_on_device, const double *bin_edges, int bins, const int32_t *win_lo,
                         const int32_t *win_hi, int hist_stride, double wl_f, int64_t n_trials, int nbr_swap,
                         uint64_t seed, uint64_t offset, int64_t *n_accept, double *e_final);

/* enter_energy_window (src/wang-landau.F90:643-741) for the first n_walkers replicas: a biased
 * walk, accept iff log(u) < -((E'-target)^2 - (E-target)^2) * inv_two_sigma_sq, until
 * lo_e[w] < E < hi_e[w] (the reference's min_e+condition / max_e-condition) or max_trials trials.
 * energies[w] returns the running energy, entered[w] = 1 on success. */
int brawl_cuda_wl_enter_window(brawl_cuda_t *h, int n_walkers, const double *target, const double *lo_e,
                               const double *hi_e, double inv_two_sigma_sq, int64_t max_trials, uint64_t seed,
                               uint64_t offset, double *energies, int32_t *entered);
/* Deterministic form for ONE walker, consuming the reference's MT19937 stream in the reference's order (two
 * rdm_site draws = 6 uniforms per iteration, + log(genrand()) when the species differ, :710-731): bit-exact with the
 * Fortran loop.  Runs iterations of its `do while` body from running energy e_start until
 *   *status = 1  the running energy satisfies lo_e < E < hi_e (:691): the caller recomputes the exact energy
 *                (brawl_cuda_total_energy, exact_order = 1) and stops, or resumes with it (resume = 0) -- the `cycle`;
 *   *status = 2  i_steps reached a multiple of `period` (= n_atoms*250, :677): the caller re-randomises the
 *                configuration (initial_setup on the same MT stream, set_config), leaves the running energy alone as the
 *                reference does, and resumes the same iteration with resume = 1;
 *   *status = 0  max_iters iterations were begun.
 * two_sigma_sq = 2*(0.0025*|energy_max - energy_min|*n_atoms/(Ry_to_eV*1000))**2, the divisor at :727-729.
 * *i_steps_io carries the reference's i_steps across calls; *iters_begun returns the iterations started by this call. */
int brawl_cuda_wl_enter_window_replay(brawl_cuda_t *h, int replica, double e_start, double target, double lo_e,
                                      double hi_e, double two_sigma_sq, int64_t period, int64_t *i_steps_io,
                                      int64_t max_iters, int resume, uint32_t *mt_state625, double *e_out, int *status,
                                      int64_t *iters_begun);
/* Device storage of the compact lattices ([n_replicas][bytes_per_replica] uint8, species 0..S-1),
 * for device-side exchange between GPUs (replica_exchange, src/wang-landau.F90:1475-1495) */
int brawl_cuda_lattice_ptr(brawl_cuda_t *h, void **dev_ptr, int64_t *bytes_per_replica);
/* exchange the configurations of two replicas on the device (same-GPU replica exchange) */
int brawl_cuda_swap_replicas(brawl_cuda_t *h, int a, int b);

/* Average a device array a[walker][len] over the walkers of each window held by this handle
 * (walkers q*wpw .. q*wpw+wpw-1 form window q) and write the mean/`divisor`-scaled sum back to
 * every walker: the on-GPU part of the allreduce + "/num_walkers" at src/wang-landau.F90:628-631. */
int brawl_cuda_wl_window_average(brawl_cuda_t *h, double *dev_array, int len, int walkers_per_window,
                                 int n_windows, double divisor);

/* ---- Wang-Landau, device-resident ------------------------------------------------------------
 * The walkers of a window live on one GPU: window q of this handle = replicas q*wpw .. q*wpw+wpw-1.  ln g[walker][bins]
 * and hist[walker][bins] (hist[i] counts bin win_lo + i) stay in HBM between calls.
 *   wl_init          allocate + zero the state (create_energy_bins' edges, :969-986)
 *   wl_set_windows   mpi_start_idx / mpi_end_idx per walker (1-based, equal inside a window); zero_hist != 0 clears hist
 *   wl_set_lng       every walker := lng[bins] (the MPI_BCAST that ends dos_combine, :1192)
 *   wl_get           what = 0: ln g, 1: hist of each window (its first walker; equal after an iterate): out[n_windows][bins]
 *   wl_iterate       one pass of the f-loop body (:214-226): sweeps for n_trials trials per walker (:539-626) from the
 *                    exact total energy (:547), then the MPI_Allreduce + "/num_walkers" of ln g and hist over each window
 *                    (:628-631) and the flatness inputs minval(hist), sum(hist)/mpi_bins per window (:222-226).  One host
 *                    synchronisation; outputs (any may be NULL): energies[n_walkers], hist_min / hist_mean[n_windows],
 *                    n_accept[n_walkers].  Fails if a walker enters with an energy outside its window. */
int brawl_cuda_wl_init(brawl_cuda_t *h, int bins, const double *bin_edges, int walkers_per_window);
int brawl_cuda_wl_set_windows(brawl_cuda_t *h, const int32_t *win_lo, const int32_t *win_hi, int zero_hist);
/* Windows that span GPUs: the windows of this handle are the SAME windows on all n_ranks ranks of its communicator
 * (brawl_cuda_comm_create first), each rank holding walkers_per_window of their walkers.  wl_iterate then divides the
 * local sums by the total number of walkers and sums ln g and hist over the ranks with ncclAllReduce -- the
 * MPI_Allreduce of src/wang-landau.F90:628-631 across GPUs (a run with fewer windows than GPUs, e.g. the one-window
 * performance/wl_input.inp, uses every GPU this way).  n_ranks = 1 (default): every window lives on one GPU. */
int brawl_cuda_wl_set_span(brawl_cuda_t *h, int n_ranks);
int brawl_cuda_wl_zero_hist(brawl_cuda_t *h);
int brawl_cuda_wl_set_lng(brawl_cuda_t *h, const double *lng);
int brawl_cuda_wl_get(brawl_cuda_t *h, int what, double *out);
int brawl_cuda_wl_iterate(brawl_cuda_t *h, double wl_f, int64_t n_trials, int nbr_swap, uint64_t seed, uint64_t offset,
                          double *energies, double *hist_min, double *hist_mean, int64_t *n_accept);

/* ---- collectives of the multi-GPU drivers (NCCL over NVLink; replaces the MPI calls of src/comms.F90 and
 * src/wang-landau.F90 listed in SURVEY 2b) ----------------------------------------------------------
 * One communicator per handle = per MPI rank / GPU.  Rank 0 obtains a 128-byte id (comm_unique_id) and hands it to the
 * other ranks by whatever the host program has (MPI_Bcast in the Fortran drivers, a file in brawl_driver,
 * torch.distributed in the Python drivers); every rank then calls comm_create.  NCCL is dlopen()ed on first use.
 *   comm_allgather       recv[n_ranks][n] := send[n] of every rank (host buffers)       -- walker energies, :1435
 *   wl_allreduce         buf[n] := sum over ranks (host buffer)                          -- converged flags :230, wl_mc_steps :244
 *   wl_allgather_lng     lng_all[n_ranks*n_windows][bins] from device memory             -- dos_combine, :1161-1192
 *   exchange_replica     swap a local configuration with one of rank `peer`, which makes the matching call
 *                        (grouped ncclSend/ncclRecv of the compact lattice)              -- replica_exchange, :1485-1495 */
int brawl_cuda_comm_unique_id(uint8_t *id128);
int brawl_cuda_comm_create(brawl_cuda_t *h, int n_ranks, int rank, const uint8_t *id128);
int brawl_cuda_comm_destroy(brawl_cuda_t *h);
int brawl_cuda_comm_allgather(brawl_cuda_t *h, const double *send, int n, double *recv);
int brawl_cuda_wl_allreduce(brawl_cuda_t *h, double *buf, int n);
int brawl_cuda_wl_allgather_lng(brawl_cuda_t *h, double *lng_all);
int brawl_cuda_exchange_replica(brawl_cuda_t *h, int replica, int peer);
/* n exchanges in one NCCL group; both sides of every pair must list it at the same position among their common pairs */
int brawl_cuda_exchange_replicas(brawl_cuda_t *h, int n, const int32_t *replica, const int32_t *peer);
/* all same-GPU swaps of one replica_exchange call in one launch: replicas a[i] <-> b[i], pairs disjoint */
int brawl_cuda_swap_replicas_batch(brawl_cuda_t *h, int n_pairs, const int32_t *a, const int32_t *b);
/* replica dst[i] := replica src[i] for n_pairs pairs in one launch (the walker clone of nested_sampling.f90:151 for a batch
 * of independent runs); destinations distinct, no destination is also a source, src[i] == dst[i] allowed (no-op) */
int brawl_cuda_copy_replicas_batch(brawl_cuda_t *h, int n_pairs, const int32_t *src, const int32_t *dst);

/* ---- nested sampling ------------------------------------------------------------------------
 * The constrained random walk of nested_sampling.f90:157-192 for a batch of walkers: walker w
 * (replica walker_ids[w]) with running energy energies[w] takes n_steps steps (site 2 redrawn
 * until species differ; accept iff E + dE < e_limit[w]).  Replay form: one walker, MT stream. */
int brawl_cuda_ns_walk_replay(brawl_cuda_t *h, int replica, double *energy, double e_limit, int64_t n_steps,
                              uint32_t *mt_state625, int64_t *n_accept);
int brawl_cuda_ns_walk(brawl_cuda_t *h, int n_walkers, const int32_t *walker_ids, double *energies,
                       const double *e_limit, int64_t n_steps, uint64_t seed, uint64_t offset, int64_t *n_accept);

#ifdef __cplusplus
}
#endif
#endif /* BRAWL_CUDA_H */
