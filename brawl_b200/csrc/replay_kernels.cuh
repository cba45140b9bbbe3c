// replay_kernels.cuh -- deterministic single-chain kernels that consume the reference's MT19937
// stream in the reference's order, so trajectories are bit-identical to the Fortran code:
//   * Metropolis   monte_carlo_step_lattice / _nbr     (src/metropolis.F90:751-891)
//   * Wang-Landau  sweeps for one walker               (src/wang-landau.F90:539-626)
//   * nested-sampling constrained walk                 (src/nested_sampling.f90:157-192)
// One warp per chain.  The chain is inherently serial (each trial's proposal depends on the RNG
// position left by the previous accept/reject), so the warp parallelises *inside* a trial: the
// 32 lanes gather the 2*Z neighbour species and look up V for the four (site, centre-species)
// combinations, then lanes 0..3 each add one combination's terms in the reference's order
// (sequential per shell, shells left to right) -- bit-exact f64 association.
#pragma once
#include "brawl_common.cuh"

struct BrwWarpScratch {
  double val[4][BRW_MAX_Z];          // [0]=site1 before, [1]=site2 before, [2]=site1 after, [3]=site2 after
  double ssum[4][BRW_MAX_SHELLS];    // per-shell partial sums of the four combinations
};
struct BrwReplaySmem {
  BrwMT mt;
  BrwWarpScratch w;
};

// neighbour-index providers for brw_warp_pair_energies
struct BrwNbrWrap {          // wrap arithmetic on the doubled grid (any lattice size)
  const BrwGeom &g;
  int x1, y1, z1, x2, y2, z2;
  __device__ __forceinline__ int n1(int k) const { return brw_nbr(g, x1, y1, z1, k); }
  __device__ __forceinline__ int n2(int k) const { return brw_nbr(g, x2, y2, z2, k); }
};
struct BrwNbrTable {         // precomputed table tab[site*ztot + k] (small lattices, shared memory)
  const unsigned short *t1, *t2;
  __device__ __forceinline__ int n1(int k) const { return t1[k]; }
  __device__ __forceinline__ int n2(int k) const { return t2[k]; }
};

// Full-warp evaluation of pair_energy before and after exchanging compact sites c1 and c2.
// Returns (identical in all lanes) before/after.
// Arithmetic: each (combination, shell) chain is a sequential f64 sum from 0.0 in the reference's
// neighbour order, done by one lane; lanes 0..3 then combine the shells left to right -- exactly the
// association of <lattice>_shellK_energy / <lattice>_energy_Nshells.  V may live in global or
// shared memory.
template <class Nbr>
__device__ __forceinline__ void brw_warp_pair_energies_t(const BrwGeom &g, const double *V, const uint8_t *lat,
                                                         const Nbr &nb, int c1, int c2, int s1, int s2,
                                                         BrwWarpScratch *w, double &before, double &after) {
  const int lane = threadIdx.x & 31;
  const int SS = g.S * g.S;
  for (int k = lane; k < g.ztot; k += 32) {
    const double *Vn = V + g.off[k][3] * SS;
    int n1 = nb.n1(k), n2 = nb.n2(k);
    int t1 = lat[n1], t2 = lat[n2];
    w->val[0][k] = Vn[t1 * g.S + s1];
    w->val[1][k] = Vn[t2 * g.S + s2];
    // after the exchange: site 1 holds s2, site 2 holds s1, and a neighbour that *is* one of the
    // two sites shows the exchanged occupant
    int u1 = n1 == c1 ? s2 : (n1 == c2 ? s1 : t1);
    int u2 = n2 == c1 ? s2 : (n2 == c2 ? s1 : t2);
    w->val[2][k] = Vn[u1 * g.S + s2];
    w->val[3][k] = Vn[u2 * g.S + s1];
  }
  __syncwarp();
  {
    const int c = lane & 3;
    const double *v = w->val[c];
    for (int n = lane >> 2; n < g.n_shells; n += 8) {
      double e = 0.0;
      const int end = g.shell_end[n];
#pragma unroll 4
      for (int k = n ? g.shell_end[n - 1] : 0; k < end; k++) e = __dadd_rn(e, v[k]);
      w->ssum[c][n] = e;
    }
  }
  __syncwarp();
  double tot = 0.0;
  if (lane < 4) {
    tot = w->ssum[lane][0];
    for (int n = 1; n < g.n_shells; n++) tot = __dadd_rn(tot, w->ssum[lane][n]);
  }
  double e0 = __shfl_sync(0xffffffffu, tot, 0), e1 = __shfl_sync(0xffffffffu, tot, 1);
  double e2 = __shfl_sync(0xffffffffu, tot, 2), e3 = __shfl_sync(0xffffffffu, tot, 3);
  before = __dadd_rn(e0, e1);
  after = __dadd_rn(e2, e3);
  __syncwarp();
}
// grid-coordinate form used by the replay kernels (`need_after` false: Wang-Landau same-species
// proposals reuse pair_unswapped, src/wang-landau.F90:561-569)
__device__ __forceinline__ void brw_warp_pair_energies(const BrwGeom &g, const double *__restrict__ V,
                                                       const uint8_t *lat, int x1, int y1, int z1, int x2,
                                                       int y2, int z2, int s1, int s2, bool need_after,
                                                       BrwWarpScratch *w, double &before, double &after) {
  BrwNbrWrap nb{g, x1, y1, z1, x2, y2, z2};
  brw_warp_pair_energies_t(g, V, lat, nb, brw_grid_to_compact(g, x1, y1, z1), brw_grid_to_compact(g, x2, y2, z2), s1, s2,
                           w, before, after);
  if (!need_after) after = before;
}

__device__ __forceinline__ void brw_mt_load(BrwMT *mt, const uint32_t *state625) {
  for (int i = threadIdx.x; i < 624; i += blockDim.x) mt->mt[i] = state625[i];
  if (threadIdx.x == 0) mt->mti = (int)state625[624];
  __syncthreads();
}
__device__ __forceinline__ void brw_mt_store(const BrwMT *mt, uint32_t *state625) {
  __syncthreads();
  for (int i = threadIdx.x; i < 624; i += blockDim.x) state625[i] = mt->mt[i];
  if (threadIdx.x == 0) state625[624] = (uint32_t)mt->mti;
}

// lane 0 draws the proposal; broadcast to the warp
__device__ __forceinline__ void brw_warp_propose(const BrwGeom &g, BrwMT *mt, int nbr_swap, int &x1, int &y1,
                                                 int &z1, int &x2, int &y2, int &z2) {
  if ((threadIdx.x & 31) == 0) {
    double u1 = brw_mt_genrand(mt), u2 = brw_mt_genrand(mt), u3 = brw_mt_genrand(mt);
    brw_random_site(g, u1, u2, u3, x1, y1, z1);
    if (nbr_swap) brw_random_nbr(g, brw_mt_genrand(mt), x1, y1, z1, x2, y2, z2);
    else {
      u1 = brw_mt_genrand(mt); u2 = brw_mt_genrand(mt); u3 = brw_mt_genrand(mt);
      brw_random_site(g, u1, u2, u3, x2, y2, z2);
    }
  }
  x1 = __shfl_sync(0xffffffffu, x1, 0); y1 = __shfl_sync(0xffffffffu, y1, 0); z1 = __shfl_sync(0xffffffffu, z1, 0);
  x2 = __shfl_sync(0xffffffffu, x2, 0); y2 = __shfl_sync(0xffffffffu, y2, 0); z2 = __shfl_sync(0xffffffffu, z2, 0);
}

// One Metropolis trial by a full warp; returns accept (1/0) in all lanes.
__device__ __forceinline__ int brw_warp_mc_step(const BrwGeom &g, const double *__restrict__ V, uint8_t *lat,
                                                BrwReplaySmem *sm, double beta, int nbr_swap) {
  int x1 = 0, y1 = 0, z1 = 0, x2 = 0, y2 = 0, z2 = 0;
  brw_warp_propose(g, &sm->mt, nbr_swap, x1, y1, z1, x2, y2, z2);
  const int c1 = brw_grid_to_compact(g, x1, y1, z1), c2 = brw_grid_to_compact(g, x2, y2, z2);
  const int s1 = lat[c1], s2 = lat[c2];
  if (s1 == s2) return 1;                                   // :774-777 (no energy, no RNG)
  double before, after;
  brw_warp_pair_energies(g, V, lat, x1, y1, z1, x2, y2, z2, s1, s2, true, &sm->w, before, after);
  const double delta_e = __dsub_rn(after, before);          // :792
  int accept = 0;
  if ((threadIdx.x & 31) == 0) {
    if (delta_e < 0.0) accept = 1;                          // :796
    else if (brw_mt_genrand(&sm->mt) < exp(-beta * delta_e)) accept = 1;   // :802
    if (accept) { lat[c1] = (uint8_t)s2; lat[c2] = (uint8_t)s1; }
  }
  accept = __shfl_sync(0xffffffffu, accept, 0);
  __syncwarp();
  return accept;
}

// Metropolis replay.  grid = 1 CTA of 256 threads: warp 0 runs the chain; when sampling is on,
// the whole CTA evaluates the per-site energies every n_sample trials and thread 0 adds them in
// reference order (exact total_energy) -- the fused form of src/metropolis.F90:348-367.
__global__ void __launch_bounds__(256) brw_metropolis_replay_kernel(BrwGeom g, const double *__restrict__ V,
                                                                    uint8_t *lat, double beta, long n_trials,
                                                                    long n_sample, int nbr_swap, uint32_t *state625,
                                                                    unsigned long long *n_accept, double *energies,
                                                                    double *site_e) {
  __shared__ BrwReplaySmem sm;
  brw_mt_load(&sm.mt, state625);
  long acc = 0;
  const long chunk = n_sample > 0 ? n_sample : n_trials;
  long done = 0, isamp = 0;
  while (done < n_trials) {
    long m = n_trials - done < chunk ? n_trials - done : chunk;
    if (threadIdx.x < 32)
      for (long t = 0; t < m; t++) acc += brw_warp_mc_step(g, V, lat, &sm, beta, nbr_swap);
    done += m;
    __syncthreads();
    if (n_sample > 0 && m == chunk) {
      for (int c = threadIdx.x; c < g.n_sites; c += blockDim.x) {
        int x, y, z;
        brw_compact_to_grid(g, c, x, y, z);
        site_e[c] = brw_site_energy(g, V, x, y, z, lat[c], BrwPlainSpec{lat});
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        double e = 0.0;
        for (int c = 0; c < g.n_sites; c++) e = __dadd_rn(e, site_e[c]);
        energies[isamp] = 0.5 * e;
      }
      isamp++;
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) *n_accept = (unsigned long long)acc;
  brw_mt_store(&sm.mt, state625);
}

// ---- Wang-Landau sweeps, one walker, MT stream (src/wang-landau.F90:539-626) -------------------
__device__ __forceinline__ int brw_bin_index(double e, double edge0, double range, int bins) {   // :515-523
  return (int)(((e - edge0) / range) * (double)bins) + 1;
}
__global__ void __launch_bounds__(32) brw_wl_replay_kernel(BrwGeom g, const double *__restrict__ V, uint8_t *lat,
                                                           double *lng, double *hist, double edge0, double range,
                                                           int bins, int win_lo, int win_hi, double wl_f,
                                                           long n_trials, int hist_every, int nbr_swap, double e_start,
                                                           uint32_t *state625, unsigned long long *n_accept,
                                                           double *e_final) {
  __shared__ BrwReplaySmem sm;
  brw_mt_load(&sm.mt, state625);
  const int lane = threadIdx.x;
  double e_unswapped = e_start, e_swapped = e_start;
  long accepted = 0;
  for (long i = 1; i <= n_trials; i++) {
    int x1 = 0, y1 = 0, z1 = 0, x2 = 0, y2 = 0, z2 = 0;
    brw_warp_propose(g, &sm.mt, nbr_swap, x1, y1, z1, x2, y2, z2);
    const int c1 = brw_grid_to_compact(g, x1, y1, z1), c2 = brw_grid_to_compact(g, x2, y2, z2);
    const int s1 = lat[c1], s2 = lat[c2];
    double pair_unswapped, pair_swapped;
    brw_warp_pair_energies(g, V, lat, x1, y1, z1, x2, y2, z2, s1, s2, s1 != s2, &sm.w, pair_unswapped, pair_swapped);
    e_swapped = e_unswapped;
    if (s1 != s2) e_swapped = __dadd_rn(__dsub_rn(e_unswapped, pair_unswapped), pair_swapped);   // :568
    int ibin = brw_bin_index(e_unswapped, edge0, range, bins), jbin = brw_bin_index(e_swapped, edge0, range, bins);
    if (lane == 0) {
      if (jbin > win_lo - 1 && jbin < win_hi + 1) {
        if (log(brw_mt_genrand(&sm.mt)) < (lng[ibin - 1] - lng[jbin - 1])) {   // :598
          accepted++;
          e_unswapped = e_swapped;
          lat[c1] = (uint8_t)s2; lat[c2] = (uint8_t)s1;
        } else jbin = ibin;
      } else jbin = ibin;
      if (hist_every > 0 && i % hist_every == 0) hist[jbin - win_lo] += 1.0;   // :605-606
      lng[jbin - 1] += wl_f;                                                   // :612 / :624
    }
    e_unswapped = __shfl_sync(0xffffffffu, e_unswapped, 0);
    __syncwarp();
  }
  if (lane == 0) { *n_accept = (unsigned long long)accepted; *e_final = e_unswapped; }
  brw_mt_store(&sm.mt, state625);
}

// ---- enter_energy_window, one walker, MT stream (src/wang-landau.F90:643-741) --------------------
// Runs iterations of the reference's `do while(.True.)` body until
//   status 1: the RUNNING energy passes the window test (:691) -- the caller recomputes the exact total energy (:692)
//             and either stops or resumes with it (`cycle`, resume = 0);
//   status 2: i_steps reached a multiple of `period` = n_atoms*250 (:677-680) -- the caller re-randomises the lattice with
//             initial_setup on the same MT stream WITHOUT touching the running energy and resumes the SAME iteration
//             (resume = 1: the window test comes next);
//   status 0: max_iters iterations begun.
// delta_e is formed with the reference's division by 2 sigma^2 (`denom`), not a multiplication by its inverse.
__global__ void __launch_bounds__(32) brw_wl_enter_replay_kernel(BrwGeom g, const double *__restrict__ V, uint8_t *lat,
                                                                 double e_start, double target, double lo, double hi,
                                                                 double denom, long period, long i_steps, long max_iters,
                                                                 int resume, uint32_t *state625, double *e_out,
                                                                 long *out3 /* status, iterations begun, i_steps */) {
  __shared__ BrwReplaySmem sm;
  brw_mt_load(&sm.mt, state625);
  const int lane = threadIdx.x;
  double e_unswapped = e_start;
  long it = 0;
  int status = 0;
  while (true) {
    if (!resume) {
      if (it >= max_iters) break;
      it++;
      i_steps++;
      if (i_steps % period == 0) { i_steps = 0; status = 2; break; }
    }
    resume = 0;
    if (e_unswapped < hi && e_unswapped > lo) { status = 1; break; }
    int x1 = 0, y1 = 0, z1 = 0, x2 = 0, y2 = 0, z2 = 0;
    brw_warp_propose(g, &sm.mt, 0, x1, y1, z1, x2, y2, z2);
    const int c1 = brw_grid_to_compact(g, x1, y1, z1), c2 = brw_grid_to_compact(g, x2, y2, z2);
    const int s1 = lat[c1], s2 = lat[c2];
    if (s1 != s2) {                                                                          // :718
      double pair_unswapped, pair_swapped;
      brw_warp_pair_energies(g, V, lat, x1, y1, z1, x2, y2, z2, s1, s2, true, &sm.w, pair_unswapped, pair_swapped);
      const double e_swapped = __dadd_rn(__dsub_rn(e_unswapped, pair_unswapped), pair_swapped);   // :724
      const double d1 = __dsub_rn(e_swapped, target), d0 = __dsub_rn(e_unswapped, target);
      const double delta_e = __ddiv_rn(__dsub_rn(__dmul_rn(d1, d1), __dmul_rn(d0, d0)), denom);   // :727-729
      if (lane == 0) {
        if (log(brw_mt_genrand(&sm.mt)) < -delta_e) {                                        // :731
          e_unswapped = e_swapped;
          lat[c1] = (uint8_t)s2; lat[c2] = (uint8_t)s1;
        }
      }
      e_unswapped = __shfl_sync(0xffffffffu, e_unswapped, 0);
      __syncwarp();
    }
  }
  if (lane == 0) { *e_out = e_unswapped; out3[0] = status; out3[1] = it; out3[2] = i_steps; }
  brw_mt_store(&sm.mt, state625);
}

// ---- nested-sampling walk, one walker, MT stream (src/nested_sampling.f90:157-192) -------------
__global__ void __launch_bounds__(32) brw_ns_replay_kernel(BrwGeom g, const double *__restrict__ V, uint8_t *lat,
                                                           double *energy_io, double e_limit, long n_steps,
                                                           uint32_t *state625, unsigned long long *n_accept) {
  __shared__ BrwReplaySmem sm;
  brw_mt_load(&sm.mt, state625);
  const int lane = threadIdx.x;
  double E = *energy_io;
  long n_acc = 0;
  for (long st = 0; st < n_steps; st++) {
    int x1 = 0, y1 = 0, z1 = 0, x2 = 0, y2 = 0, z2 = 0, c1 = 0, c2 = 0, s1 = 0, s2 = 0;
    if (lane == 0) {
      double u1 = brw_mt_genrand(&sm.mt), u2 = brw_mt_genrand(&sm.mt), u3 = brw_mt_genrand(&sm.mt);
      brw_random_site(g, u1, u2, u3, x1, y1, z1);
      c1 = brw_grid_to_compact(g, x1, y1, z1); s1 = lat[c1];
      do {                                                        // :162-173
        u1 = brw_mt_genrand(&sm.mt); u2 = brw_mt_genrand(&sm.mt); u3 = brw_mt_genrand(&sm.mt);
        brw_random_site(g, u1, u2, u3, x2, y2, z2);
        c2 = brw_grid_to_compact(g, x2, y2, z2); s2 = lat[c2];
      } while (s1 == s2);
    }
    x1 = __shfl_sync(0xffffffffu, x1, 0); y1 = __shfl_sync(0xffffffffu, y1, 0); z1 = __shfl_sync(0xffffffffu, z1, 0);
    x2 = __shfl_sync(0xffffffffu, x2, 0); y2 = __shfl_sync(0xffffffffu, y2, 0); z2 = __shfl_sync(0xffffffffu, z2, 0);
    c1 = __shfl_sync(0xffffffffu, c1, 0); c2 = __shfl_sync(0xffffffffu, c2, 0);
    s1 = __shfl_sync(0xffffffffu, s1, 0); s2 = __shfl_sync(0xffffffffu, s2, 0);
    double before, after;
    brw_warp_pair_energies(g, V, lat, x1, y1, z1, x2, y2, z2, s1, s2, true, &sm.w, before, after);
    const double delta_e = __dsub_rn(after, before);
    if (__dadd_rn(E, delta_e) < e_limit) {                        // :183-186
      E = __dadd_rn(E, delta_e);
      n_acc++;
      if (lane == 0) { lat[c1] = (uint8_t)s2; lat[c2] = (uint8_t)s1; }
    }
    __syncwarp();
  }
  if (lane == 0) { *energy_io = E; *n_accept = (unsigned long long)n_acc; }
  brw_mt_store(&sm.mt, state625);
}
