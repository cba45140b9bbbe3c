// walker_kernels.cuh -- production (Philox) kernels for batches of independent walkers/replicas on
// small lattices: ONE WARP PER WALKER, each an exact sequential chain with the reference's proposal
// distribution; the walker's lattice is staged in shared memory for the whole call.  The energy change of a trial
// is either the reference's f64 association (brw_warp_pair_energies: per-shell sequential sums through a
// shared-memory scratch, ~1000 cycles of dependent latency; handles with dE_mode 0, and lattices without a neighbour
// table) or -- the default -- a lane-parallel sum: every lane adds the V differences of its <= ceil(Z/32)
// neighbours of both sites, then a 5-stage xor butterfly (same value in all lanes, fixed order => deterministic):
// ~300 cycles.  It differs from the reference association by f64 rounding only (a walker's running energy stays
// within 1e-11 Ry of the exact total energy over 10^4 trials, tested); bit-exact trajectories are the job of the
// replay kernels.
//   Metropolis : src/metropolis.F90:751-891 (k-loop :348-354), lattices too small for boxes
//   WL         : src/wang-landau.F90:539-626 (sweeps), :515-523 (bin_index), :643-741 (enter_energy_window)
//   NS         : src/nested_sampling.f90:157-192
// RNG: Philox4x32-10, counter = (trial, stream tag, walker, offset), key = seed.  Lane 0 draws.
#pragma once
#include "brawl_common.cuh"
#include "replay_kernels.cuh"   // brw_warp_pair_energies, brw_bin_index

#define BRW_WALKER_WARPS 4          // warps (= walkers) per CTA
#define BRW_WALKER_SMEM_SITES 4096  // lattices up to this many sites are staged in shared memory
#define BRW_WALKER_TAB_BYTES 65536  // neighbour-index table is used when n_sites*ztot*2 fits in this

// Dynamic shared-memory layout of the walker kernels (host mirror: brw_walker_layout):
//   [ V : S*S*n_shells f64 ][ nbr table : n_sites*ztot u16 (optional) ]
//   then per warp: [ lattice : n_sites bytes, padded to 8 (optional) ][ extra f64 words ]
struct BrwWalkerLayout {
  int use_tab, staged;
  int fast_delta;              // lane-parallel dE (needs the table); 0: reference association
  int v_bytes, tab_bytes, ksh_bytes, lat_bytes, extra_bytes;   // ksh = shell index per neighbour; extra = per-warp f64 scratch (WL: ln g + hist)
  __host__ __device__ size_t total() const { return (size_t)v_bytes + tab_bytes + ksh_bytes + (size_t)BRW_WALKER_WARPS * (lat_bytes + extra_bytes); }
};
static inline BrwWalkerLayout brw_walker_layout(const BrwGeom &g, int extra_doubles) {
  BrwWalkerLayout L;
  L.v_bytes = g.S * g.S * g.n_shells * 8;
  L.staged = g.n_sites <= BRW_WALKER_SMEM_SITES;
  L.use_tab = L.staged && (size_t)g.n_sites * g.ztot * 2 <= BRW_WALKER_TAB_BYTES && g.n_sites <= 65535;
  L.tab_bytes = L.use_tab ? ((g.n_sites * g.ztot * 2 + 7) & ~7) : 0;
  L.ksh_bytes = L.use_tab ? ((g.ztot + 7) & ~7) : 0;
  L.fast_delta = L.use_tab;
  L.lat_bytes = L.staged ? ((g.n_sites + 7) & ~7) : 0;
  L.extra_bytes = extra_doubles * 8;
  return L;
}

struct BrwWalkerCtx {
  uint8_t *L;                  // the walker's lattice (shared-memory copy, or global if too large)
  uint8_t *G;                  // its home in global memory
  const double *V;             // V_ex in shared memory
  const unsigned short *tab;   // neighbour table or nullptr
  const unsigned char *ksh;    // shell index of neighbour k (shared memory; with the table)
  bool fast;                   // lane-parallel dE
  double *extra;               // per-warp f64 scratch
  bool staged;
  uint32_t Ls;                 // shared-memory address of the staged lattice (explicit ld.shared: the generic L
                               // pointer may also be global, so plain loads through it compile to slower generic LD)
};
__device__ __forceinline__ int brw_lds_u8_at(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return (int)v;
}
// species at compact site i of the walker's lattice
__device__ __forceinline__ int brw_walker_site(const BrwWalkerCtx &c, int i) {
  return c.staged ? brw_lds_u8_at(c.Ls + (uint32_t)i) : (int)c.L[i];
}

// CTA-wide set-up: V and the neighbour table are built cooperatively, each warp stages its lattice.
__device__ __forceinline__ BrwWalkerCtx brw_walker_begin(const BrwGeom &g, const BrwWalkerLayout &lay,
                                                         const double *__restrict__ Vg, uint8_t *lat_global,
                                                         unsigned char *smem, bool valid) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *Vs = reinterpret_cast<double *>(smem);
  unsigned short *tab = reinterpret_cast<unsigned short *>(smem + lay.v_bytes);
  for (int i = threadIdx.x; i < g.S * g.S * g.n_shells; i += blockDim.x) Vs[i] = Vg[i];
  if (lay.use_tab)
    for (int i = threadIdx.x; i < g.n_sites * g.ztot; i += blockDim.x) {
      int c = i / g.ztot, k = i - c * g.ztot, x, y, z;
      brw_compact_to_grid(g, c, x, y, z);
      tab[i] = (unsigned short)brw_nbr(g, x, y, z, k);
    }
  unsigned char *ksh = smem + lay.v_bytes + lay.tab_bytes;
  if (lay.use_tab)
    for (int k = threadIdx.x; k < g.ztot; k += blockDim.x) ksh[k] = (unsigned char)g.off[k][3];
  BrwWalkerCtx c;
  unsigned char *mine = smem + lay.v_bytes + lay.tab_bytes + lay.ksh_bytes + (size_t)warp * (lay.lat_bytes + lay.extra_bytes);
  c.ksh = ksh;
  c.fast = lay.use_tab && lay.fast_delta;
  c.G = lat_global;
  c.V = Vs;
  c.tab = lay.use_tab ? tab : nullptr;
  c.staged = lay.staged;
  c.L = lay.staged ? mine : lat_global;
  c.Ls = lay.staged ? (uint32_t)__cvta_generic_to_shared(mine) : 0u;
  c.extra = reinterpret_cast<double *>(mine + lay.lat_bytes);
  if (lay.staged && valid)
    for (int i = lane; i < g.n_sites; i += 32) c.L[i] = lat_global[i];
  __syncthreads();
  return c;
}
__device__ __forceinline__ void brw_walker_end(const BrwGeom &g, const BrwWalkerCtx &c) {
  if (c.staged) {
    __syncwarp();
    for (int i = (threadIdx.x & 31); i < g.n_sites; i += 32) c.G[i] = c.L[i];
  }
}
__device__ __forceinline__ void brw_walker_pair(const BrwGeom &g, const BrwWalkerCtx &c, int x1, int y1, int z1, int x2,
                                                int y2, int z2, int c1, int c2, int s1, int s2, BrwWarpScratch *w,
                                                double &before, double &after) {
  if (c.fast) {
    // lane-parallel: d = sum_k [V_k(u1k, s2) - V_k(t1k, s1)] + [V_k(u2k, s1) - V_k(t2k, s2)], t = occupant of
    // neighbour k before the exchange, u = after it (a neighbour that IS one of the two sites shows the exchanged
    // occupant: src/metropolis.F90:783-792 swaps, then recomputes)
    const int lane = threadIdx.x & 31, S = g.S, SS = S * S;
    const unsigned short *t1 = c.tab + c1 * g.ztot, *t2 = c.tab + c2 * g.ztot;
    double d = 0.0;
#pragma unroll 2
    for (int k = lane; k < g.ztot; k += 32) {
      const double *Vn = c.V + c.ksh[k] * SS;
      const int n1 = t1[k], n2 = t2[k];
      const int a = brw_lds_u8_at(c.Ls + n1), b = brw_lds_u8_at(c.Ls + n2);      // fast => table => staged
      const int u1 = n1 == c1 ? s2 : (n1 == c2 ? s1 : a);
      const int u2 = n2 == c1 ? s2 : (n2 == c2 ? s1 : b);
      d += (Vn[u1 * S + s2] - Vn[a * S + s1]) + (Vn[u2 * S + s1] - Vn[b * S + s2]);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) d += __shfl_xor_sync(0xffffffffu, d, m);
    before = 0.0; after = d;
  } else if (c.tab) {
    BrwNbrTable nb{c.tab + c1 * g.ztot, c.tab + c2 * g.ztot};
    brw_warp_pair_energies_t(g, c.V, c.L, nb, c1, c2, s1, s2, w, before, after);
  } else {
    BrwNbrWrap nb{g, x1, y1, z1, x2, y2, z2};
    brw_warp_pair_energies_t(g, c.V, c.L, nb, c1, c2, s1, s2, w, before, after);
  }
}
struct BrwProposal;
__device__ __forceinline__ void brw_walker_pair_lane(const BrwGeom &g, const BrwWalkerCtx &c, const BrwProposal &mine, int src,
                                                     int c1, int c2, int s1, int s2, BrwWarpScratch *w, double &before,
                                                     double &after);
// Proposals are counter-based (Philox counter = trial index), so they can be generated 32 trials at a
// time: lane l computes the proposal of trial (base + l) -- two Philox blocks, the reference's
// random_site / random_nbr arithmetic -- and each trial then broadcasts its lane's values.  Same
// counters, same values as a per-trial draw; ~30x fewer RNG instructions per trial.
struct BrwProposal {
  int x1, y1, z1, x2, y2, z2;
  int c1, c2;                // compact indices of the two sites (all the table-driven paths need)
  uint32_t spare;            // 4th word of the second block: the uniform of the accept test
};
__device__ __forceinline__ BrwProposal brw_propose_philox(const BrwGeom &g, int nbr_swap, long t, uint32_t tag,
                                                          uint32_t walker, uint32_t off_lo, uint32_t k0, uint32_t k1) {
  BrwProposal p;
  const uint32_t t_lo = (uint32_t)t, t_hi = (uint32_t)(t >> 32) ^ tag;
  BrwPhilox4 r1 = brw_philox(t_lo, t_hi ^ 0x10000000u, walker, off_lo, k0, k1);
  BrwPhilox4 r2 = brw_philox(t_lo, t_hi ^ 0x20000000u, walker, off_lo, k0, k1);
  brw_random_site(g, brw_u01(r1.x), brw_u01(r1.y), brw_u01(r1.z), p.x1, p.y1, p.z1);
  if (nbr_swap) brw_random_nbr(g, brw_u01(r1.w), p.x1, p.y1, p.z1, p.x2, p.y2, p.z2);
  else brw_random_site(g, brw_u01(r2.x), brw_u01(r2.y), brw_u01(r2.z), p.x2, p.y2, p.z2);
  p.spare = r2.w;
  p.c1 = brw_grid_to_compact(g, p.x1, p.y1, p.z1);
  p.c2 = brw_grid_to_compact(g, p.x2, p.y2, p.z2);
  return p;
}
__device__ __forceinline__ BrwProposal brw_proposal_from_lane(const BrwProposal &mine, int src) {
  BrwProposal p;
  p.x1 = __shfl_sync(0xffffffffu, mine.x1, src); p.y1 = __shfl_sync(0xffffffffu, mine.y1, src);
  p.z1 = __shfl_sync(0xffffffffu, mine.z1, src); p.x2 = __shfl_sync(0xffffffffu, mine.x2, src);
  p.y2 = __shfl_sync(0xffffffffu, mine.y2, src); p.z2 = __shfl_sync(0xffffffffu, mine.z2, src);
  p.spare = __shfl_sync(0xffffffffu, mine.spare, src);
  p.c1 = __shfl_sync(0xffffffffu, mine.c1, src); p.c2 = __shfl_sync(0xffffffffu, mine.c2, src);
  return p;
}
// the three values every trial needs (compact sites + the spare random word); the grid coordinates are fetched only
// by the path without a neighbour table (brw_walker_pair_lane)
__device__ __forceinline__ void brw_proposal_sites_from_lane(const BrwProposal &mine, int src, int &c1, int &c2, uint32_t &spare) {
  c1 = __shfl_sync(0xffffffffu, mine.c1, src); c2 = __shfl_sync(0xffffffffu, mine.c2, src);
  spare = __shfl_sync(0xffffffffu, mine.spare, src);
}

__device__ __forceinline__ void brw_walker_pair_lane(const BrwGeom &g, const BrwWalkerCtx &c, const BrwProposal &mine, int src,
                                                     int c1, int c2, int s1, int s2, BrwWarpScratch *w, double &before,
                                                     double &after) {
  if (c.tab) brw_walker_pair(g, c, 0, 0, 0, 0, 0, 0, c1, c2, s1, s2, w, before, after);      // coordinates unused with the table
  else {
    const BrwProposal pr = brw_proposal_from_lane(mine, src);
    brw_walker_pair(g, c, pr.x1, pr.y1, pr.z1, pr.x2, pr.y2, pr.z2, c1, c2, s1, s2, w, before, after);
  }
}

// ---- Metropolis chains -------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * BRW_WALKER_WARPS) brw_chain_metropolis_kernel(
    BrwGeom g, BrwWalkerLayout lay, const double *__restrict__ V, uint8_t *lat, const double *__restrict__ beta,
    int n_rep, long n_trials, int nbr_swap, uint32_t k0, uint32_t k1, uint32_t off_lo, uint32_t off_hi,
    unsigned long long *att_out, unsigned long long *acc_out, double *dE_out) {
  __shared__ BrwWarpScratch scratch[BRW_WALKER_WARPS];
  extern __shared__ __align__(16) unsigned char dsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * BRW_WALKER_WARPS + warp;
  const bool valid = r < n_rep;
  BrwWalkerCtx c = brw_walker_begin(g, lay, V, lat + (long)(valid ? r : 0) * g.n_sites, dsm, valid);
  if (!valid) return;
  const double b = beta[r];
  unsigned long long acc = 0;
  double dsum = 0.0;
  BrwProposal mine = {};
  for (long t = 0; t < n_trials; t++) {
    if ((t & 31) == 0) mine = brw_propose_philox(g, nbr_swap, t + lane, 0u, (uint32_t)r, off_lo, k0, k1 ^ off_hi);
    int c1, c2; uint32_t w;
    brw_proposal_sites_from_lane(mine, (int)(t & 31), c1, c2, w);
    const int s1 = brw_walker_site(c, c1), s2 = brw_walker_site(c, c2);
    if (s1 == s2) { acc++; continue; }                        // src/metropolis.F90:774-777
    double before, after;
    brw_walker_pair_lane(g, c, mine, (int)(t & 31), c1, c2, s1, s2, &scratch[warp], before, after);
    const double dE = __dsub_rn(after, before);
    bool accept = dE < 0.0;
    if (!accept) accept = brw_u01(w) < exp(-b * dE);
    if (accept) {
      if (lane == 0) { c.L[c1] = (uint8_t)s2; c.L[c2] = (uint8_t)s1; }
      acc++; dsum += dE;
    }
    __syncwarp();
  }
  brw_walker_end(g, c);
  if (lane == 0) { att_out[r] += (unsigned long long)n_trials; acc_out[r] += acc; dE_out[r] += dsum; }
}

// bin_index (src/wang-landau.F90:515-523) without the f64 division on the common path: t = (e - e0) * (1/range) * bins
// differs from the reference's ((e - e0) / range) * bins by < 1e-12, so the truncated integer can differ only if t lies
// within that distance of an integer -- then (|t - rint(t)| < 1e-6) the reference expression is evaluated instead.
// Same result as brw_bin_index for every input; ~100 cycles less dependent latency per trial.
__device__ __forceinline__ int brw_bin_index_nodiv(double e, double edge0, double range, double inv_range, int bins) {
  const double t = ((e - edge0) * inv_range) * (double)bins;
  if (fabs(t - rint(t)) < 1e-6 || !(fabs(t) < 1e9)) return brw_bin_index(e, edge0, range, bins);
  return (int)t + 1;
}

// ---- Wang-Landau sweeps ---------------------------------------------------------------------------
// ln g (all bins) and the window's hist slice of each walker are staged in shared memory for the call.
__global__ void __launch_bounds__(32 * BRW_WALKER_WARPS) brw_wl_walker_kernel(
    BrwGeom g, BrwWalkerLayout lay, const double *__restrict__ V, uint8_t *lat, double *lng, double *hist, double edge0,
    double range, int bins, const int *__restrict__ win_lo, const int *__restrict__ win_hi, int hist_stride, double wl_f,
    long n_trials, int hist_every, int nbr_swap, uint32_t k0, uint32_t k1, uint32_t off_lo, uint32_t off_hi,
    int n_walkers, double *e_io, unsigned long long *n_accept, int *bad_walker = nullptr) {
  __shared__ BrwWarpScratch scratch[BRW_WALKER_WARPS];
  extern __shared__ __align__(16) unsigned char dsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * BRW_WALKER_WARPS + warp;
  const bool valid = w < n_walkers;
  BrwWalkerCtx c = brw_walker_begin(g, lay, V, lat + (long)(valid ? w : 0) * g.n_sites, dsm, valid);
  if (!valid) return;
  double *g_lng = lng + (long)w * bins, *g_hist = hist + (long)w * hist_stride;
  const int lo = win_lo[w], hi = win_hi[w], nh = hi - lo + 1;
  if (bad_walker) {
    // device-resident callers cannot check the start energies on the host: a walker outside its window (the reference
    // would index wl_logdos out of bounds) is reported through the flag and left untouched
    const double e0 = e_io[w];
    const int ib = brw_bin_index(e0, edge0, range, bins);
    if (!(e0 == e0) || ib < lo || ib > hi) {
      if (lane == 0) { atomicCAS(bad_walker, 0, w + 1); n_accept[w] = 0; }
      return;
    }
  }
  double *my_lng = c.extra, *my_hist = c.extra + bins;
  for (int i = lane; i < bins; i += 32) my_lng[i] = g_lng[i];
  for (int i = lane; i < nh; i += 32) my_hist[i] = g_hist[i];
  __syncwarp();
  double e_unswapped = e_io[w], e_swapped;
  unsigned long long accepted = 0;
  BrwProposal mine = {};
  double my_logu = 0.0;
  long hist_left = hist_every > 0 ? hist_every : -1;           // trials until the next histogram sample (i % hist_every == 0)
  int ibin = brw_bin_index(e_unswapped, edge0, range, bins);
  const double inv_range = 1.0 / range;
  for (long i = 1; i <= n_trials; i++) {
    const int slot = (int)((i - 1) & 31);
    if (slot == 0) {
      mine = brw_propose_philox(g, nbr_swap, i + lane, 0x01000000u, (uint32_t)w, off_lo, k0, k1 ^ off_hi);
      my_logu = log(brw_u01(mine.spare));      // u == 0 gives -inf: accepted, as in the reference
    }
    int c1, c2; uint32_t spare_unused;
    brw_proposal_sites_from_lane(mine, slot, c1, c2, spare_unused);
    const double logu = __shfl_sync(0xffffffffu, my_logu, slot);
    const int s1 = brw_walker_site(c, c1), s2 = brw_walker_site(c, c2);
    e_swapped = e_unswapped;
    if (s1 != s2) {
      double pair_unswapped, pair_swapped;
      brw_walker_pair_lane(g, c, mine, slot, c1, c2, s1, s2, &scratch[warp], pair_unswapped, pair_swapped);
      e_swapped = __dadd_rn(__dsub_rn(e_unswapped, pair_unswapped), pair_swapped);      // :568
    }
    // bin of the current state is carried from trial to trial (same value as recomputing it: src/wang-landau.F90:571);
    // a same-species proposal leaves the energy, hence the bin, unchanged
    const int jbin_new = s1 != s2 ? brw_bin_index_nodiv(e_swapped, edge0, range, inv_range, bins) : ibin;
    int jbin = jbin_new;
    // decision on lane 0 (it owns ln g / hist)
    int acc = 0;
    if (lane == 0) {
      if (jbin > lo - 1 && jbin < hi + 1) {
        if (logu < (my_lng[ibin - 1] - my_lng[jbin - 1])) {                              // :598
          acc = 1;
          c.L[c1] = (uint8_t)s2; c.L[c2] = (uint8_t)s1;
        } else jbin = ibin;
      } else jbin = ibin;
      if (--hist_left == 0) { my_hist[jbin - lo] += 1.0; hist_left = hist_every; }        // :605-606
      my_lng[jbin - 1] += wl_f;                                                          // :612 / :624
    }
    acc = __shfl_sync(0xffffffffu, acc, 0);
    if (acc) { accepted++; e_unswapped = e_swapped; ibin = jbin_new; }
    __syncwarp();
  }
  for (int i = lane; i < bins; i += 32) g_lng[i] = my_lng[i];
  for (int i = lane; i < nh; i += 32) g_hist[i] = my_hist[i];
  brw_walker_end(g, c);
  if (lane == 0) { e_io[w] = e_unswapped; n_accept[w] = accepted; }
}

// enter_energy_window (src/wang-landau.F90:643-741): biased walk towards the window centre,
// accept iff log(u) < -((E'-E*)^2 - (E-E*)^2) * inv_two_sigma_sq, until lo_e < E < hi_e (the
// reference's min_e+condition / max_e-condition) or max_trials.  entered[w] = 1 on success.
__global__ void __launch_bounds__(32 * BRW_WALKER_WARPS) brw_wl_enter_window_kernel(
    BrwGeom g, BrwWalkerLayout lay, const double *__restrict__ V, uint8_t *lat, const double *__restrict__ target,
    const double *__restrict__ lo_e, const double *__restrict__ hi_e, double inv_two_sigma_sq, long max_trials,
    uint32_t k0, uint32_t k1, uint32_t off_lo, uint32_t off_hi, int n_walkers, double *e_io, int *entered) {
  __shared__ BrwWarpScratch scratch[BRW_WALKER_WARPS];
  extern __shared__ __align__(16) unsigned char dsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * BRW_WALKER_WARPS + warp;
  const bool valid = w < n_walkers;
  BrwWalkerCtx c = brw_walker_begin(g, lay, V, lat + (long)(valid ? w : 0) * g.n_sites, dsm, valid);
  if (!valid) return;
  const double tgt = target[w], lo = lo_e[w], hi = hi_e[w];
  double e = e_io[w];
  int ok = (e < hi && e > lo) ? 1 : 0;
  BrwProposal mine = {};
  for (long i = 0; i < max_trials && !ok; i++) {
    if ((i & 31) == 0) mine = brw_propose_philox(g, 0, i + lane, 0x02000000u, (uint32_t)w, off_lo, k0, k1 ^ off_hi);
    int c1, c2; uint32_t u;
    brw_proposal_sites_from_lane(mine, (int)(i & 31), c1, c2, u);
    const int s1 = brw_walker_site(c, c1), s2 = brw_walker_site(c, c2);
    if (s1 != s2) {                                                                       // :710
      double pu, ps;
      brw_walker_pair_lane(g, c, mine, (int)(i & 31), c1, c2, s1, s2, &scratch[warp], pu, ps);
      const double e_new = __dadd_rn(__dsub_rn(e, pu), ps);                                // :721
      const double delta = ((e_new - tgt) * (e_new - tgt) - (e - tgt) * (e - tgt)) * inv_two_sigma_sq;   // :725-727
      if (log(brw_u01(u)) < -delta) {                                                      // :729
        e = e_new;
        if (lane == 0) { c.L[c1] = (uint8_t)s2; c.L[c2] = (uint8_t)s1; }
      }
      __syncwarp();
    }
    ok = (e < hi && e > lo) ? 1 : 0;
  }
  brw_walker_end(g, c);
  if (lane == 0) { e_io[w] = e; entered[w] = ok; }
}

// intra-window average of ln g and hist over the walkers of each window held on this GPU
// (the on-device part of the MPI_Allreduce/num_walkers at src/wang-landau.F90:628-631).
// Walkers of window q are w = q*wpw .. q*wpw+wpw-1.  Sum in walker order, then divide.
__global__ void brw_wl_window_average_kernel(double *a, int len, int wpw, int n_windows, double divisor) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_windows * len) return;
  int q = idx / len, b = idx - q * len;
  double s = 0.0;
  for (int k = 0; k < wpw; k++) s += a[((long)(q * wpw + k)) * len + b];
  s = s / divisor;
  for (int k = 0; k < wpw; k++) a[((long)(q * wpw + k)) * len + b] = s;
}

// ---- nested-sampling walks ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * BRW_WALKER_WARPS) brw_ns_walker_kernel(
    BrwGeom g, BrwWalkerLayout lay, const double *__restrict__ V, uint8_t *lat, const int *__restrict__ walker_ids,
    double *energies,
    const double *__restrict__ e_limit, long n_steps, uint32_t k0, uint32_t k1, uint32_t off_lo, uint32_t off_hi,
    int n_walkers, unsigned long long *n_accept) {
  __shared__ BrwWarpScratch scratch[BRW_WALKER_WARPS];
  extern __shared__ __align__(16) unsigned char dsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * BRW_WALKER_WARPS + warp;
  const bool valid = w < n_walkers;
  BrwWalkerCtx c = brw_walker_begin(g, lay, V, lat + (long)(valid ? walker_ids[w] : 0) * g.n_sites, dsm, valid);
  if (!valid) return;
  double E = energies[w];
  const double lim = e_limit[w];
  unsigned long long n_acc = 0;
  int bx1 = 0, by1 = 0, bz1 = 0, bx2 = 0, by2 = 0, bz2 = 0;    // this lane's proposal for step (base + lane)
  for (long st = 0; st < n_steps; st++) {
    if ((st & 31) == 0) {                                       // 32 steps' first proposals at once (same counters as per-step draws)
      const long t = st + lane;
      BrwPhilox4 r1 = brw_philox((uint32_t)t, (uint32_t)(t >> 32) ^ 0x50000000u, (uint32_t)w, off_lo, k0, k1 ^ off_hi);
      BrwPhilox4 r2 = brw_philox((uint32_t)t, ((uint32_t)(t >> 32) & 0xFFFFu) ^ 0x60000000u, (uint32_t)w, off_lo, k0, k1 ^ off_hi);
      brw_random_site(g, brw_u01(r1.x), brw_u01(r1.y), brw_u01(r1.z), bx1, by1, bz1);
      brw_random_site(g, brw_u01(r2.x), brw_u01(r2.y), brw_u01(r2.z), bx2, by2, bz2);
    }
    const int src = (int)(st & 31);
    int x1 = __shfl_sync(0xffffffffu, bx1, src), y1 = __shfl_sync(0xffffffffu, by1, src), z1 = __shfl_sync(0xffffffffu, bz1, src);
    int x2 = __shfl_sync(0xffffffffu, bx2, src), y2 = __shfl_sync(0xffffffffu, by2, src), z2 = __shfl_sync(0xffffffffu, bz2, src);
    int c1 = brw_grid_to_compact(g, x1, y1, z1), c2 = brw_grid_to_compact(g, x2, y2, z2);
    int s1 = c.L[c1], s2 = c.L[c2];
    if (s1 == s2) {                                             // :162-173 redraw site 2 until species differ
      if (lane == 0) {
        uint32_t tries = 1;
        do {
          BrwPhilox4 r2 = brw_philox((uint32_t)st, ((uint32_t)(st >> 32) & 0xFFFFu) ^ 0x60000000u ^ (tries << 16),
                                     (uint32_t)w, off_lo, k0, k1 ^ off_hi);
          brw_random_site(g, brw_u01(r2.x), brw_u01(r2.y), brw_u01(r2.z), x2, y2, z2);
          c2 = brw_grid_to_compact(g, x2, y2, z2);
          s2 = c.L[c2];
          tries++;
        } while (s1 == s2 && tries < 4096u);
      }
      x2 = __shfl_sync(0xffffffffu, x2, 0); y2 = __shfl_sync(0xffffffffu, y2, 0); z2 = __shfl_sync(0xffffffffu, z2, 0);
      c2 = __shfl_sync(0xffffffffu, c2, 0); s2 = __shfl_sync(0xffffffffu, s2, 0);
    }
    if (s1 == s2) continue;                                    // single-species lattice: nothing to do
    double before, after;
    brw_walker_pair(g, c, x1, y1, z1, x2, y2, z2, c1, c2, s1, s2, &scratch[warp], before, after);
    const double dE = __dsub_rn(after, before);
    if (__dadd_rn(E, dE) < lim) {                              // :183-186
      E = __dadd_rn(E, dE);
      n_acc++;
      if (lane == 0) { c.L[c1] = (uint8_t)s2; c.L[c2] = (uint8_t)s1; }
    }
    __syncwarp();
  }
  brw_walker_end(g, c);
  if (lane == 0) { energies[w] = E; n_accept[w] = n_acc; }
}
