// brawl_common.cuh -- geometry, exact-order energy arithmetic and RNGs shared by all kernels of
// libbrawl_cuda.so (sm_100a).  Product code: must not include anything from oracle/.
//
// HBM layout.  The reference keeps config(1,2n1,2n2,2n3) with only 1/4 (bcc) or 1/2 (fcc) of the
// cells occupied (src/initialise.F90:547-610).  On the device each replica is stored compact,
// sites only, species 0..S-1:  lat[(z*cy + yc)*cx + xc]
//      bcc: xc = x>>1, yc = y>>1   (sites have x = y = z mod 2)      cx=n1  cy=n2   cz=2n3
//      fcc: xc = x>>1, yc = y      (sites have x+y+z even)           cx=n1  cy=2n2  cz=2n3
//      sc : xc = x,    yc = y                                        cx=2n1 cy=2n2  cz=2n3
// The compact linear order equals the reference's z/y/x traversal order over occupied cells, so
// "sum in compact order" == total_energy's summation order (src/bw_hamiltonian.f90:67-76).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define BRW_MAX_Z 168       // bcc, 10 shells
#define BRW_MAX_SHELLS 10
#define BRW_MAX_SPECIES 16

struct BrwGeom {            // POD, passed by value as a kernel parameter (< 1 KiB)
  int lattice, S, n_shells, ztot;
  int gx, gy, gz;           // doubled grid extents 2n
  int wx, wy, wz;           // neighbour wrap moduli (2n; n for simple cubic -- reference quirk,
                            // src/bw_hamiltonian.f90:1944-1949)
  int cx, cy, cz;           // compact extents
  int xs, ys;               // compact shifts: xc = x >> xs, yc = y >> ys
  int n_sites;              // cx*cy*cz == n_atoms
  int shell_end[BRW_MAX_SHELLS];      // cumulative neighbour counts per shell
  signed char off[BRW_MAX_Z][4];      // dx, dy, dz, shell -- reference summation order
};

// ---- geometry --------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void brw_compact_to_grid(const BrwGeom &g, int c, int &x, int &y, int &z) {
  int xc = c % g.cx, t = c / g.cx, yc = t % g.cy;
  z = t / g.cy;
  if (g.lattice == 1) { x = 2 * xc + (z & 1); y = 2 * yc + (z & 1); }
  else if (g.lattice == 2) { y = yc; x = 2 * xc + ((y + z) & 1); }
  else { x = xc; y = yc; }
}
__host__ __device__ __forceinline__ bool brw_is_site(const BrwGeom &g, int x, int y, int z) {
  if (g.lattice == 1) return ((x ^ z) & 1) == 0 && ((y ^ z) & 1) == 0;
  if (g.lattice == 2) return ((x + y + z) & 1) == 0;
  return true;
}
__host__ __device__ __forceinline__ int brw_grid_to_compact(const BrwGeom &g, int x, int y, int z) {
  return (z * g.cy + (y >> g.ys)) * g.cx + (x >> g.xs);
}
__host__ __device__ __forceinline__ int brw_wrap(int t, int w) {
  while (t < 0) t += w;
  while (t >= w) t -= w;
  return t;
}
// compact index of neighbour k of grid site (x,y,z)
__device__ __forceinline__ int brw_nbr(const BrwGeom &g, int x, int y, int z, int k) {
  int nx = brw_wrap(x + g.off[k][0], g.wx);
  int ny = brw_wrap(y + g.off[k][1], g.wy);
  int nz = brw_wrap(z + g.off[k][2], g.wz);
  return brw_grid_to_compact(g, nx, ny, nz);
}

// ---- exact-order energies --------------------------------------------------------------------
// setup%nbr_energy: per shell a sequential f64 sum from 0.0 in listed-neighbour order
// (e.g. src/bw_hamiltonian.f90:171-173), shells combined left to right ((s1+s2)+s3)+...
// (e.g. :1014-1017).  `spec(c)` returns the 0-based species at compact index c, which lets
// callers evaluate "after the swap" energies without touching memory.  Adds only (no FMA risk);
// __dadd_rn pins round-to-nearest regardless of compiler flags.
template <class Spec>
__device__ __forceinline__ double brw_site_energy(const BrwGeom &g, const double *__restrict__ V, int x, int y,
                                                  int z, int centre, Spec spec) {
  double tot = 0.0;
  int k = 0;
  const int SS = g.S * g.S;
  for (int n = 0; n < g.n_shells; n++) {
    double e = 0.0;
    const double *Vn = V + n * SS + centre;
    const int end = g.shell_end[n];
    for (; k < end; k++) {
      int s = spec(brw_nbr(g, x, y, z, k));
      e = __dadd_rn(e, __ldg(Vn + s * g.S));
    }
    tot = (n == 0) ? e : __dadd_rn(tot, e);
  }
  return tot;
}

struct BrwPlainSpec {
  const uint8_t *lat;
  __device__ __forceinline__ int operator()(int c) const { return lat[c]; }
};
// species as they would be after exchanging the occupants of compact sites c1 <-> c2
struct BrwSwappedSpec {
  const uint8_t *lat;
  int c1, c2, s1, s2;   // s1 = species currently at c1, s2 = at c2
  __device__ __forceinline__ int operator()(int c) const { return c == c1 ? s2 : (c == c2 ? s1 : lat[c]); }
};

// pair_energy before / after exchanging c1 <-> c2 (src/bw_hamiltonian.f90:99-114 evaluated around
// pair_swap as in src/metropolis.F90:783-789).  One thread.
__device__ __forceinline__ void brw_pair_energies(const BrwGeom &g, const double *__restrict__ V,
                                                  const uint8_t *lat, int c1, int c2, double &before,
                                                  double &after) {
  int x1, y1, z1, x2, y2, z2;
  brw_compact_to_grid(g, c1, x1, y1, z1);
  brw_compact_to_grid(g, c2, x2, y2, z2);
  int s1 = lat[c1], s2 = lat[c2];
  BrwPlainSpec p{lat};
  before = __dadd_rn(brw_site_energy(g, V, x1, y1, z1, s1, p), brw_site_energy(g, V, x2, y2, z2, s2, p));
  BrwSwappedSpec q{lat, c1, c2, s1, s2};
  after = __dadd_rn(brw_site_energy(g, V, x1, y1, z1, q(c1), q), brw_site_energy(g, V, x2, y2, z2, q(c2), q));
}

// ---- MT19937 (reference stream; src/mt19937ar.c:82-125) ---------------------------------------
struct BrwMT {
  uint32_t mt[624];
  int mti;
};
__device__ __forceinline__ uint32_t brw_mt_int32(BrwMT *g) {
  if (g->mti >= 624) {
    uint32_t *mt = g->mt;
    for (int k = 0; k < 624; k++) {
      int k1 = k + 1 == 624 ? 0 : k + 1;
      int km = k + 397 >= 624 ? k + 397 - 624 : k + 397;
      uint32_t y = (mt[k] & 0x80000000u) | (mt[k1] & 0x7fffffffu);
      mt[k] = mt[km] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->mti = 0;
  }
  uint32_t y = g->mt[g->mti++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
__device__ __forceinline__ double brw_mt_genrand(BrwMT *g) { return brw_mt_int32(g) * (1.0 / 4294967296.0); }

// ---- Philox4x32-10 (Salmon et al., SC'11), counter-based production streams --------------------
struct BrwPhilox4 { uint32_t x, y, z, w; };
__host__ __device__ __forceinline__ BrwPhilox4 brw_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return BrwPhilox4{c0, c1, c2, c3};
}
// uniform in [0,1) with the same 32-bit resolution as genrand() (src/mt19937ar.c:121-125)
__host__ __device__ __forceinline__ double brw_u01(uint32_t r) { return r * (1.0 / 4294967296.0); }
// unbiased-enough integer in [0,n): multiply-high (n << 2^32 everywhere it is used)
__host__ __device__ __forceinline__ uint32_t brw_below(uint32_t r, uint32_t n) {
  return (uint32_t)(((uint64_t)r * n) >> 32);
}

// ---- site proposal (src/random_site.f90:74-225), 0-based grid coordinates ----------------------
// u1,u2,u3 are consumed in the reference's order: z first, then x, then y (sc: x, y, z).
__device__ __forceinline__ void brw_random_site(const BrwGeom &g, double u1, double u2, double u3, int &x,
                                                int &y, int &z) {
  if (g.lattice == 1) {
    int z1 = (int)floor(2.0 * u1 * (double)(g.gz >> 1)) + 1;
    x = 2 * (int)floor(u2 * (double)(g.gx >> 1)) + 2 - (z1 & 1) - 1;
    y = 2 * (int)floor(u3 * (double)(g.gy >> 1)) + 2 - (z1 & 1) - 1;
    z = z1 - 1;
  } else if (g.lattice == 2) {
    int z1 = (int)floor(2.0 * u1 * (double)(g.gz >> 1)) + 1;
    int x1 = (int)floor(2.0 * u2 * (double)(g.gx >> 1)) + 1;
    int y1 = 2 * (int)floor(u3 * (double)(g.gy >> 1)) + 1 + ((x1 - (z1 & 1)) & 1);
    x = x1 - 1; y = y1 - 1; z = z1 - 1;
  } else {
    x = (int)floor(u1 * 2.0 * (double)(g.gx >> 1));
    y = (int)floor(u2 * 2.0 * (double)(g.gy >> 1));
    z = (int)floor(u3 * 2.0 * (double)(g.gz >> 1));
  }
}
// first-shell neighbour tables of random_site.f90:29-61 (order differs from the energy tables)
static __device__ __constant__ signed char brw_rnbr_sc[6][3] = {{0,0,1},{0,1,0},{1,0,0},{0,0,-1},{0,-1,0},{-1,0,0}};
static __device__ __constant__ signed char brw_rnbr_bcc[8][3] = {{1,1,1},{1,1,-1},{1,-1,1},{1,-1,-1},{-1,1,1},{-1,1,-1},{-1,-1,1},{-1,-1,-1}};
static __device__ __constant__ signed char brw_rnbr_fcc[12][3] = {{0,1,1},{0,1,-1},{0,-1,1},{0,-1,-1},{1,1,0},{1,-1,0},{1,0,1},{1,0,-1},{-1,1,0},{-1,-1,0},{-1,0,1},{-1,0,-1}};
__device__ __forceinline__ void brw_random_nbr(const BrwGeom &g, double u, int x, int y, int z, int &nx, int &ny,
                                               int &nz) {
  int z1 = g.lattice == 0 ? 6 : g.lattice == 1 ? 8 : 12;
  int n = (int)floor((double)z1 * u);
  const signed char *t = g.lattice == 0 ? brw_rnbr_sc[n] : g.lattice == 1 ? brw_rnbr_bcc[n] : brw_rnbr_fcc[n];
  nx = brw_wrap(x + t[0], g.gx);
  ny = brw_wrap(y + t[1], g.gy);
  nz = brw_wrap(z + t[2], g.gz);
}

// ---- host-side context (defined in brawl_api.cu) ------------------------------------------------
struct brawl_cuda_ctx {
  int device;
  cudaStream_t stream, own_stream;
  BrwGeom g;
  int n_replicas;
  int64_t grid_cells;          // 8*n1*n2*n3
  uint8_t *d_lat;              // [n_replicas][n_sites] compact species 0..S-1
  double *d_V;                 // V_ex, Fortran order
  int8_t *d_stage;             // staging for reference-layout grids (grown on demand)
  size_t stage_bytes;
  double *d_scratch;           // per-site energies / partial sums (grown on demand)
  size_t scratch_bytes;
  int *d_flag;                 // device error flag
  double *hV;                  // host copy of V_ex
  double *d_beta;              // [n_replicas]
  void *d_small;               // small results / scalars buffer (grown on demand)
  size_t small_bytes;
  // production Metropolis plans: [0] whole-lattice swaps, [1] neighbour swaps (BrwPlan*)
  void *mc_plan[2];
  int last_plan;
  int tune_box[3], tune_steps, disable_fast, cubic_period_only, last_launches;
  int dE_mode;                 // 0: reference association for every trial; 1: screened, byte lattice; 2: screened, word lattice (default)
  int word_split;              // 1 (default): the planner may pick the two-warp-group (SPLIT) word kernels
  int word_epoch;              // steps per epoch of the site-energy-caching kernels (epoch_metropolis.cuh): 4 (default), 8, 2, or 0 = off
  int byte_layout;             // 1: never use the word-lattice kernels / dense decomposition (test hook, A/B comparisons)
  uint32_t *d_order;           // [n_replicas][S][n_sites] occupancy counts (store_state), allocated on first use
  void *wl;                    // BrwWlState*: device-resident Wang-Landau ln g / histograms (wl_resident.inc)
  void *comm;                  // BrwComm*: NCCL communicator of the multi-GPU drivers (wl_resident.inc)
};

int brw_fail(const char *fmt, ...);              // sets last error, returns 1
int brw_cuda_check(cudaError_t e, const char *what);
int brw_ensure_scratch(brawl_cuda_ctx *h, size_t bytes);
int brw_ensure_stage(brawl_cuda_ctx *h, size_t bytes);
#define BRW_CUDA(x) do { if (brw_cuda_check((x), #x)) return 1; } while (0)
#define BRW_LAUNCH_CHECK(what) do { if (brw_cuda_check(cudaGetLastError(), what)) return 1; } while (0)
