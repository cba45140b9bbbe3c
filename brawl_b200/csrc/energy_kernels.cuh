// energy_kernels.cuh -- Hamiltonian kernels: per-site nbr_energy, full-lattice energy (exact
// reference order or deterministic tree), per-swap dE batch, SRO pair counts, and the
// reference-grid <-> compact-lattice converters.  Included by brawl_cuda.cu.
#pragma once
#include "brawl_common.cuh"

// ---- layout conversion -----------------------------------------------------------------------
// grid[z][y][x] int8 (reference layout, species 1..S, 0 off-site) -> compact uint8 species 0..S-1.
// flag bit0: bad species on a site; bit1: non-zero off-site cell.
__global__ void brw_pack_kernel(BrwGeom g, const int8_t *__restrict__ grid, uint8_t *__restrict__ lat, int n_rep,
                                int *flag) {
  const long cells = (long)g.gx * g.gy * g.gz;
  const long total = cells * n_rep;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i / cells, c = i - r * cells;
    int x = (int)(c % g.gx), y = (int)((c / g.gx) % g.gy), z = (int)(c / ((long)g.gx * g.gy));
    int v = grid[i];
    if (brw_is_site(g, x, y, z)) {
      if (v < 1 || v > g.S) { atomicOr(flag, 1); v = 1; }
      lat[r * g.n_sites + brw_grid_to_compact(g, x, y, z)] = (uint8_t)(v - 1);
    } else if (v != 0) atomicOr(flag, 2);
  }
}
__global__ void brw_unpack_kernel(BrwGeom g, const uint8_t *__restrict__ lat, int8_t *__restrict__ grid, int n_rep) {
  const long cells = (long)g.gx * g.gy * g.gz;
  const long total = cells * n_rep;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i / cells, c = i - r * cells;
    int x = (int)(c % g.gx), y = (int)((c / g.gx) % g.gy), z = (int)(c / ((long)g.gx * g.gy));
    grid[i] = brw_is_site(g, x, y, z) ? (int8_t)(lat[r * g.n_sites + brw_grid_to_compact(g, x, y, z)] + 1) : (int8_t)0;
  }
}

// 16 grid cells (one uint4) per thread, for grids whose x extent is a multiple of 16: the 16 cells lie in one row
// (y, z), whose site parity is thread-uniform, and map to 8 (bcc, fcc) or 16 (sc) consecutive compact bytes.
__global__ void __launch_bounds__(256) brw_pack16_kernel(BrwGeom g, const int8_t *__restrict__ grid,
                                                         uint8_t *__restrict__ lat, int n_rep, int *flag) {
  const long cells = (long)g.gx * g.gy * g.gz, chunks = cells / 16, total = chunks * n_rep;
  const int cpr = g.gx / 16;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / chunks, c = i - r * chunks;
    const int k = (int)(c % cpr), row = (int)(c / cpr), y = row % g.gy, z = row / g.gy;
    const uint4 v = reinterpret_cast<const uint4 *>(grid)[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    // x-parity of the sites of this row (-1: the row holds no sites)
    const int par = g.lattice == 1 ? (((y ^ z) & 1) ? -1 : (z & 1)) : g.lattice == 2 ? ((y + z) & 1) : 0;
    int bad = 0;
    uint32_t o[4] = {0, 0, 0, 0};
    int no = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const int cell = (int)(int8_t)((w[j >> 2] >> (8 * (j & 3))) & 255u);
      const bool site = g.lattice == 0 ? true : (par >= 0 && (j & 1) == par);
      if (site) {
        int sp = cell;
        if (sp < 1 || sp > g.S) { bad |= 1; sp = 1; }
        o[no >> 2] |= (uint32_t)(sp - 1) << (8 * (no & 3));
        no++;
      } else if (cell != 0) bad |= 2;
    }
    if (bad) atomicOr(flag, bad);
    if (g.lattice == 0) reinterpret_cast<uint4 *>(lat + r * g.n_sites + ((long)row * g.cx + 16 * k))[0] = make_uint4(o[0], o[1], o[2], o[3]);
    else if (par >= 0) reinterpret_cast<uint2 *>(lat + r * g.n_sites + ((long)(z * g.cy + (y >> g.ys)) * g.cx + 8 * k))[0] = make_uint2(o[0], o[1]);
  }
}
__global__ void __launch_bounds__(256) brw_unpack16_kernel(BrwGeom g, const uint8_t *__restrict__ lat,
                                                           int8_t *__restrict__ grid, int n_rep) {
  const long cells = (long)g.gx * g.gy * g.gz, chunks = cells / 16, total = chunks * n_rep;
  const int cpr = g.gx / 16;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / chunks, c = i - r * chunks;
    const int k = (int)(c % cpr), row = (int)(c / cpr), y = row % g.gy, z = row / g.gy;
    const int par = g.lattice == 1 ? (((y ^ z) & 1) ? -1 : (z & 1)) : g.lattice == 2 ? ((y + z) & 1) : 0;
    uint32_t in[4] = {0, 0, 0, 0};
    if (g.lattice == 0) {
      const uint4 v = reinterpret_cast<const uint4 *>(lat + r * g.n_sites + ((long)row * g.cx + 16 * k))[0];
      in[0] = v.x; in[1] = v.y; in[2] = v.z; in[3] = v.w;
    } else if (par >= 0) {
      const uint2 v = reinterpret_cast<const uint2 *>(lat + r * g.n_sites + ((long)(z * g.cy + (y >> g.ys)) * g.cx + 8 * k))[0];
      in[0] = v.x; in[1] = v.y;
    }
    uint32_t w[4] = {0, 0, 0, 0};
    int ni = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const bool site = g.lattice == 0 ? true : (par >= 0 && (j & 1) == par);
      if (site) {
        w[j >> 2] |= (((in[ni >> 2] >> (8 * (ni & 3))) & 255u) + 1u) << (8 * (j & 3));
        ni++;
      }
    }
    reinterpret_cast<uint4 *>(grid)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---- per-site energies -----------------------------------------------------------------------
// out[r*n_sites + c] = nbr_energy of compact site c of replica r.  One thread per site.
// HBM-bound in principle (1 B read + 8 B written per site); gathers hit L1/L2.
__global__ void brw_site_energy_kernel(BrwGeom g, const double *__restrict__ V, const uint8_t *__restrict__ lat,
                                       double *__restrict__ out, int n_rep) {
  const long total = (long)g.n_sites * n_rep;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long r = i / g.n_sites;
    int c = (int)(i - r * g.n_sites);
    const uint8_t *L = lat + r * g.n_sites;
    int x, y, z;
    brw_compact_to_grid(g, c, x, y, z);
    out[i] = brw_site_energy(g, V, x, y, z, L[c], BrwPlainSpec{L});
  }
}
// scatter compact per-site energies to the reference grid shape (0.0 off-site)
__global__ void brw_site_energy_to_grid_kernel(BrwGeom g, const double *__restrict__ e, double *__restrict__ out) {
  const long cells = (long)g.gx * g.gy * g.gz;
  for (long c = blockIdx.x * (long)blockDim.x + threadIdx.x; c < cells; c += (long)gridDim.x * blockDim.x) {
    int x = (int)(c % g.gx), y = (int)((c / g.gx) % g.gy), z = (int)(c / ((long)g.gx * g.gy));
    out[c] = brw_is_site(g, x, y, z) ? e[brw_grid_to_compact(g, x, y, z)] : 0.0;
  }
}

// ---- ordered sum: bit-exact total_energy -----------------------------------------------------
// One CTA per replica.  f64 addition is not associative, so the reference's sequential z/y/x
// accumulation (src/bw_hamiltonian.f90:67-79) is reproduced literally: the CTA stages chunks of
// the per-site energies in shared memory (coalesced) and thread 0 adds them in compact order.
#define BRW_OSUM_CHUNK 2048
__global__ void __launch_bounds__(256) brw_ordered_sum_kernel(const double *__restrict__ e, long n, double *__restrict__ out) {
  __shared__ double buf[2][BRW_OSUM_CHUNK];
  const double *src = e + (long)blockIdx.x * n;
  double acc = 0.0;
  long nchunks = (n + BRW_OSUM_CHUNK - 1) / BRW_OSUM_CHUNK;
  // prefetch chunk 0
  for (int i = threadIdx.x; i < BRW_OSUM_CHUNK && i < n; i += blockDim.x) buf[0][i] = src[i];
  __syncthreads();
  for (long ch = 0; ch < nchunks; ch++) {
    int cur = ch & 1;
    long base = ch * BRW_OSUM_CHUNK, nb = base + BRW_OSUM_CHUNK;
    if (threadIdx.x == 0) {
      int m = (int)((n - base) < BRW_OSUM_CHUNK ? (n - base) : BRW_OSUM_CHUNK);
      const double *b = buf[cur];
#pragma unroll 8
      for (int i = 0; i < m; i++) acc = __dadd_rn(acc, b[i]);
    } else if (nb < n) {
      // the other 255 threads fetch the next chunk meanwhile
      for (int i = threadIdx.x - 1; i < BRW_OSUM_CHUNK && nb + i < n; i += blockDim.x - 1) buf[cur ^ 1][i] = src[nb + i];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = 0.5 * acc;   // :79
}

// ---- tree sum: fast total energy (deterministic, not reference order) -------------------------
// Fused site-energy + block reduction; partial[r*nblk + b]; finished by brw_tree_final_kernel.
// One CTA per compact row (fixed y,z): the row's neighbour rows are CTA-uniform, V sits in shared
// memory, and wraps are single conditional adds (loop fallback only for boxes smaller than the reach).
__device__ __forceinline__ int brw_wrap1(int t, int w) {
  t += (t < 0) ? w : 0;
  t -= (t >= w) ? w : 0;
  if (t < 0 || t >= w) t = brw_wrap(t, w);
  return t;
}
__global__ void __launch_bounds__(128) brw_energy_partial_kernel(BrwGeom g, const double *__restrict__ V,
                                                                 const uint8_t *__restrict__ lat,
                                                                 double *__restrict__ partial, int nblk) {
  __shared__ double red[4];
  extern __shared__ double Vs[];
  const int r = blockIdx.y;
  const uint8_t *L = lat + (long)r * g.n_sites;
  for (int i = threadIdx.x; i < g.S * g.S * g.n_shells; i += blockDim.x) Vs[i] = V[i];
  __syncthreads();
  const int SS = g.S * g.S;
  double acc = 0.0;
  const int n_rows = g.cy * g.cz;
  for (int row = blockIdx.x; row < n_rows; row += nblk) {
    const int yc = row % g.cy, z = row / g.cy;
    const int y = g.lattice == 1 ? 2 * yc + (z & 1) : yc;
    const int xpar = g.lattice == 1 ? (z & 1) : (g.lattice == 2 ? ((y + z) & 1) : 0);
    for (int xc = threadIdx.x; xc < g.cx; xc += blockDim.x) {
      const int x = g.lattice == 0 ? xc : 2 * xc + xpar;
      const int centre = L[row * g.cx + xc];
      double tot = 0.0;
      int k = 0;
      for (int n = 0; n < g.n_shells; n++) {
        double e = 0.0;
        const double *Vn = Vs + n * SS + centre;
        const int end = g.shell_end[n];
        for (; k < end; k++) {
          const int nx = brw_wrap1(x + g.off[k][0], g.wx), ny = brw_wrap1(y + g.off[k][1], g.wy),
                    nz = brw_wrap1(z + g.off[k][2], g.wz);
          const int s = L[(nz * g.cy + (ny >> g.ys)) * g.cx + (nx >> g.xs)];
          e = __dadd_rn(e, Vn[s * g.S]);
        }
        tot = (n == 0) ? e : __dadd_rn(tot, e);
      }
      acc += tot;
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += red[w];
    partial[(long)r * nblk + blockIdx.x] = s;
  }
}
// one CTA of 128 threads per replica: strided partial sums, then a fixed shared-memory tree (deterministic)
__global__ void __launch_bounds__(128) brw_tree_final_kernel(const double *__restrict__ partial, int nblk,
                                                             double *__restrict__ out, int n_rep) {
  __shared__ double red[128];
  const int r = blockIdx.x;
  if (r >= n_rep) return;
  double s = 0.0;
  for (int b = threadIdx.x; b < nblk; b += 128) s += partial[(long)r * nblk + b];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[r] = 0.5 * red[0];
}

// ---- per-swap dE batch -----------------------------------------------------------------------
// idx are flat indices on the reference grid; dE = pair_energy(after) - pair_energy(before).
__global__ void brw_pair_dE_kernel(BrwGeom g, const double *__restrict__ V, const uint8_t *__restrict__ lat, long n,
                                   const int32_t *__restrict__ i1, const int32_t *__restrict__ i2,
                                   double *__restrict__ dE, int *flag) {
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
    int a = i1[t], b = i2[t];
    int xa = a % g.gx, ya = (a / g.gx) % g.gy, za = a / (g.gx * g.gy);
    int xb = b % g.gx, yb = (b / g.gx) % g.gy, zb = b / (g.gx * g.gy);
    if (a < 0 || b < 0 || za >= g.gz || zb >= g.gz || !brw_is_site(g, xa, ya, za) || !brw_is_site(g, xb, yb, zb)) {
      atomicOr(flag, 4); dE[t] = 0.0; continue;
    }
    double before, after;
    brw_pair_energies(g, V, lat, brw_grid_to_compact(g, xa, ya, za), brw_grid_to_compact(g, xb, yb, zb), before, after);
    dE[t] = __dsub_rn(after, before);
  }
}

// ---- SRO pair counts (radial_densities, src/analytics.f90:293-404) ----------------------------
// sro_off[k] = (dx,dy,dz,l): every raw cube offset the reference's scan visits whose distance
// matches WC shell l (built on the host; l = 0 is the site itself).  Integer counts, so the
// result is exact; rho = cnt / species_count is formed by the caller in f64.
__global__ void __launch_bounds__(256) brw_radial_counts_kernel(BrwGeom g, const uint8_t *__restrict__ lat,
                                                                const int4 *__restrict__ sro_off, int n_off,
                                                                int wc_range, unsigned long long *__restrict__ cnt,
                                                                unsigned long long *__restrict__ species_count) {
  extern __shared__ unsigned int hist[];   // [wc_range*S*S] + [S]
  const int nh = wc_range * g.S * g.S + g.S;
  // batched form: blockIdx.y = replica; its counts follow the previous replica's [wc_range*S*S + S] block
  lat += (size_t)blockIdx.y * g.n_sites;
  cnt += (size_t)blockIdx.y * nh;
  species_count += (size_t)blockIdx.y * nh;
  for (int i = threadIdx.x; i < nh; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  unsigned int *sc = hist + wc_range * g.S * g.S;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < g.n_sites; c += gridDim.x * blockDim.x) {
    int x, y, z;
    brw_compact_to_grid(g, c, x, y, z);
    int si = lat[c];
    atomicAdd(&sc[si], 1u);
    for (int k = 0; k < n_off; k++) {
      int4 o = sro_off[k];
      int nx = brw_wrap(x + o.x, g.gx), ny = brw_wrap(y + o.y, g.gy), nz = brw_wrap(z + o.z, g.gz);
      int sj = lat[brw_grid_to_compact(g, nx, ny, nz)];
      atomicAdd(&hist[(o.w * g.S + sj) * g.S + si], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < wc_range * g.S * g.S; i += blockDim.x)
    if (hist[i]) atomicAdd(&cnt[i], (unsigned long long)hist[i]);
  for (int i = threadIdx.x; i < g.S; i += blockDim.x)
    if (sc[i]) atomicAdd(&species_count[i], (unsigned long long)sc[i]);
}
