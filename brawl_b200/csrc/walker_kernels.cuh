// walker_kernels.cuh -- production (Philox) Wang-Landau sweeps and nested-sampling walks for
// batches of independent walkers on small lattices: one thread per walker, each an exact
// sequential chain with the reference's proposal distribution and f64 energy association.
//   WL : src/wang-landau.F90:539-626 (sweeps), :515-523 (bin_index)
//   NS : src/nested_sampling.f90:157-192
#pragma once
#include "brawl_common.cuh"
#include "replay_kernels.cuh"   // brw_bin_index

__global__ void brw_wl_walker_kernel(BrwGeom g, const double *__restrict__ V, uint8_t *lat, double *lng, double *hist,
                                     double edge0, double range, int bins, const int *__restrict__ win_lo,
                                     const int *__restrict__ win_hi, int hist_stride, double wl_f, long n_trials,
                                     int hist_every, int nbr_swap, uint32_t k0, uint32_t k1, uint32_t off_lo,
                                     uint32_t off_hi, int n_walkers, double *e_io, unsigned long long *n_accept) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_walkers) return;
  uint8_t *L = lat + (long)w * g.n_sites;
  double *my_lng = lng + (long)w * bins, *my_hist = hist + (long)w * hist_stride;
  const int lo = win_lo[w], hi = win_hi[w];
  double e_unswapped = e_io[w], e_swapped;
  unsigned long long accepted = 0;
  for (long i = 1; i <= n_trials; i++) {
    BrwPhilox4 r1 = brw_philox((uint32_t)i, (uint32_t)(i >> 32) ^ 0x30000000u, (uint32_t)w, off_lo, k0, k1 ^ off_hi);
    BrwPhilox4 r2 = brw_philox((uint32_t)i, (uint32_t)(i >> 32) ^ 0x40000000u, (uint32_t)w, off_lo, k0, k1 ^ off_hi);
    int x1, y1, z1, x2, y2, z2;
    brw_random_site(g, brw_u01(r1.x), brw_u01(r1.y), brw_u01(r1.z), x1, y1, z1);
    if (nbr_swap) brw_random_nbr(g, brw_u01(r1.w), x1, y1, z1, x2, y2, z2);
    else brw_random_site(g, brw_u01(r2.x), brw_u01(r2.y), brw_u01(r2.z), x2, y2, z2);
    const int c1 = brw_grid_to_compact(g, x1, y1, z1), c2 = brw_grid_to_compact(g, x2, y2, z2);
    const int s1 = L[c1], s2 = L[c2];
    e_swapped = e_unswapped;
    if (s1 != s2) {
      double pair_unswapped, pair_swapped;
      brw_pair_energies(g, V, L, c1, c2, pair_unswapped, pair_swapped);
      e_swapped = __dadd_rn(__dsub_rn(e_unswapped, pair_unswapped), pair_swapped);    // :568
    }
    int ibin = brw_bin_index(e_unswapped, edge0, range, bins), jbin = brw_bin_index(e_swapped, edge0, range, bins);
    if (jbin > lo - 1 && jbin < hi + 1) {
      // u == 0 gives log = -inf: accepted, as in the reference
      if (log(brw_u01(r2.w)) < (my_lng[ibin - 1] - my_lng[jbin - 1])) {               // :598
        accepted++;
        e_unswapped = e_swapped;
        L[c1] = (uint8_t)s2; L[c2] = (uint8_t)s1;
      } else jbin = ibin;
    } else jbin = ibin;
    if (hist_every > 0 && i % hist_every == 0) my_hist[jbin - lo] += 1.0;             // :605-606
    my_lng[jbin - 1] += wl_f;                                                         // :612 / :624
  }
  e_io[w] = e_unswapped;
  n_accept[w] = accepted;
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

__global__ void brw_ns_walker_kernel(BrwGeom g, const double *__restrict__ V, uint8_t *lat,
                                     const int *__restrict__ walker_ids, double *energies,
                                     const double *__restrict__ e_limit, long n_steps, uint32_t k0, uint32_t k1,
                                     uint32_t off_lo, uint32_t off_hi, int n_walkers, unsigned long long *n_accept) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_walkers) return;
  uint8_t *L = lat + (long)walker_ids[w] * g.n_sites;
  double E = energies[w];
  const double lim = e_limit[w];
  unsigned long long n_acc = 0;
  for (long st = 0; st < n_steps; st++) {
    BrwPhilox4 r1 = brw_philox((uint32_t)st, (uint32_t)(st >> 32) ^ 0x50000000u, (uint32_t)w, off_lo, k0, k1 ^ off_hi);
    int x1, y1, z1, x2, y2, z2;
    brw_random_site(g, brw_u01(r1.x), brw_u01(r1.y), brw_u01(r1.z), x1, y1, z1);
    const int c1 = brw_grid_to_compact(g, x1, y1, z1);
    const int s1 = L[c1];
    int c2, s2;
    uint32_t tries = 0;
    do {                                                     // :162-173 redraw until species differ
      BrwPhilox4 r2 = brw_philox((uint32_t)st, ((uint32_t)(st >> 32) & 0xFFFFu) ^ 0x60000000u ^ (tries << 16),
                                 (uint32_t)w, off_lo, k0, k1 ^ off_hi);
      brw_random_site(g, brw_u01(r2.x), brw_u01(r2.y), brw_u01(r2.z), x2, y2, z2);
      c2 = brw_grid_to_compact(g, x2, y2, z2);
      s2 = L[c2];
      tries++;
    } while (s1 == s2 && tries < 4096u);
    if (s1 == s2) continue;                                  // single-species lattice: nothing to do
    double before, after;
    brw_pair_energies(g, V, L, c1, c2, before, after);
    const double dE = __dsub_rn(after, before);
    if (__dadd_rn(E, dE) < lim) {                            // :183-186
      E = __dadd_rn(E, dE);
      n_acc++;
      L[c1] = (uint8_t)s2; L[c2] = (uint8_t)s1;
    }
  }
  energies[w] = E;
  n_accept[w] = n_acc;
}
