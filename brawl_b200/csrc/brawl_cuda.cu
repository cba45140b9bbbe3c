// brawl_cuda.cu -- C ABI (include/brawl_cuda.h) of libbrawl_cuda.so.  The kernel headers are included here; the
// byte-lattice epoch kernels are instantiated in byte_epoch_kernels.cu (compiled in parallel; no device linking).
// Product code: there is NO CPU fallback -- every entry point needs a working CUDA device.
#include <array>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <cub/device/device_radix_sort.cuh>

#include "../../include/brawl_cuda.h"
#include "brawl_common.cuh"
#define BRW_TABLE_QUAL static constexpr
#include "shell_tables.inc"
#include "energy_kernels.cuh"
#include "replay_kernels.cuh"
#include "tile_metropolis.cuh"
#include "word_metropolis.cuh"
#include "epoch_metropolis.cuh"
#include "epoch_byte_metropolis.cuh"     // BrwByteEpochT (the kernels are instantiated in byte_epoch_kernels.cu)
#include "walker_kernels.cuh"

// ---- error state ---------------------------------------------------------------------------------
static thread_local std::string g_err = "";
int brw_fail(const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
int brw_cuda_check(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return 0;
  return brw_fail("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
}
extern "C" const char *brawl_cuda_last_error(void) { return g_err.c_str(); }
extern "C" int brawl_cuda_version(void) { return 1; }
extern "C" int brawl_cuda_device_count(int *count) {
  int n = 0;
  BRW_CUDA(cudaGetDeviceCount(&n));
  if (count) *count = n;
  if (n < 1) return brw_fail("no CUDA device visible");
  return 0;
}

__global__ void brw_philox_test_kernel(const uint32_t *c, const uint32_t *k, uint32_t *o) {
  BrwPhilox4 r = brw_philox(c[0], c[1], c[2], c[3], k[0], k[1]);
  o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
}
extern "C" int brawl_cuda_philox4x32(int device, const uint32_t *c, const uint32_t *k, uint32_t *o) {
  if (!c || !k || !o) return brw_fail("null argument");
  int nd = 0;
  if (brawl_cuda_device_count(&nd)) return 1;
  if (device < 0 || device >= nd) return brw_fail("device %d not available", device);
  BRW_CUDA(cudaSetDevice(device));
  uint32_t *d = nullptr;
  BRW_CUDA(cudaMalloc(&d, 10 * sizeof(uint32_t)));
  cudaMemcpy(d, c, 16, cudaMemcpyHostToDevice);
  cudaMemcpy(d + 4, k, 8, cudaMemcpyHostToDevice);
  brw_philox_test_kernel<<<1, 1>>>(d, d + 4, d + 6);
  cudaError_t e = cudaMemcpy(o, d + 6, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return brw_cuda_check(e, "philox self-test");
}

int brw_ensure_scratch(brawl_cuda_ctx *h, size_t bytes) {
  if (h->scratch_bytes >= bytes) return 0;
  if (h->d_scratch) BRW_CUDA(cudaFree(h->d_scratch));
  h->d_scratch = nullptr; h->scratch_bytes = 0;
  BRW_CUDA(cudaMalloc(&h->d_scratch, bytes));
  h->scratch_bytes = bytes;
  return 0;
}
int brw_ensure_stage(brawl_cuda_ctx *h, size_t bytes) {
  if (h->stage_bytes >= bytes) return 0;
  if (h->d_stage) BRW_CUDA(cudaFree(h->d_stage));
  h->d_stage = nullptr; h->stage_bytes = 0;
  BRW_CUDA(cudaMalloc(&h->d_stage, bytes));
  h->stage_bytes = bytes;
  return 0;
}
static int grid_for(long n, int block, int cap = 148 * 16) {
  long b = (n + block - 1) / block;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
#define BRW_ENTER(h)                                                      \
  do {                                                                    \
    if (!(h)) return brw_fail("null handle");                             \
    BRW_CUDA(cudaSetDevice((h)->device));                                 \
  } while (0)
#define BRW_REPLICA(h, r)                                                 \
  do { if ((r) < 0 || (r) >= (h)->n_replicas) return brw_fail("replica %d out of range [0,%d)", (r), (h)->n_replicas); } while (0)

// ---- handle ----------------------------------------------------------------------------------------
extern "C" int brawl_cuda_create(int lattice, int n1, int n2, int n3, int S, int n_shells, const double *V,
                                 int device, int n_replicas, brawl_cuda_t **out) {
  if (!out) return brw_fail("null handle pointer");
  *out = nullptr;
  if (lattice < 0 || lattice > 2) return brw_fail("Lattice type not yet implemented!");             // initialise.F90:238-240
  int maxs = lattice == 1 ? BRW_BCC_MAX_SHELLS : lattice == 2 ? BRW_FCC_MAX_SHELLS : BRW_SC_MAX_SHELLS;
  if (n_shells < 1 || n_shells > maxs) return brw_fail("Unsupported number of shells");              // initialise.F90:166-169 etc.
  if (n1 < 1 || n2 < 1 || n3 < 1) return brw_fail("lattice extents must be positive");
  if (S < 1 || S > BRW_MAX_SPECIES) return brw_fail("n_species must be in 1..%d", BRW_MAX_SPECIES);
  if (!V) return brw_fail("null V_ex");
  if (n_replicas < 1) return brw_fail("n_replicas must be >= 1");
  if ((long)n1 * n2 * n3 * 8 > (1L << 31) - 1) return brw_fail("lattice too large for 32-bit site indices");
  int ndev = 0;
  if (brawl_cuda_device_count(&ndev)) return 1;
  if (device < 0 || device >= ndev) return brw_fail("device %d not available (%d visible)", device, ndev);
  BRW_CUDA(cudaSetDevice(device));

  brawl_cuda_ctx *h = new brawl_cuda_ctx();
  memset(h, 0, sizeof *h);
  h->device = device;
  BrwGeom &g = h->g;
  g.lattice = lattice; g.S = S; g.n_shells = n_shells;
  g.gx = 2 * n1; g.gy = 2 * n2; g.gz = 2 * n3;
  if (lattice == 0) { g.wx = n1; g.wy = n2; g.wz = n3; g.cx = g.gx; g.cy = g.gy; g.xs = 0; g.ys = 0; }
  else { g.wx = g.gx; g.wy = g.gy; g.wz = g.gz; g.cx = n1; g.xs = 1; g.cy = lattice == 1 ? n2 : g.gy; g.ys = lattice == 1 ? 1 : 0; }
  g.cz = g.gz;
  g.n_sites = g.cx * g.cy * g.cz;
  const signed char(*tab)[3] = lattice == 1 ? brw_bcc_off : lattice == 2 ? brw_fcc_off : brw_sc_off;
  const int *st = lattice == 1 ? brw_bcc_start : lattice == 2 ? brw_fcc_start : brw_sc_start;
  const int *ct = lattice == 1 ? brw_bcc_count : lattice == 2 ? brw_fcc_count : brw_sc_count;
  int k = 0;
  for (int n = 0; n < n_shells; n++) {
    for (int j = 0; j < ct[n]; j++, k++) {
      g.off[k][0] = tab[st[n] + j][0]; g.off[k][1] = tab[st[n] + j][1]; g.off[k][2] = tab[st[n] + j][2];
      g.off[k][3] = (signed char)n;
    }
    g.shell_end[n] = k;
  }
  g.ztot = k;
  h->n_replicas = n_replicas;
  h->grid_cells = (int64_t)g.gx * g.gy * g.gz;
  h->tune_box[0] = h->tune_box[1] = h->tune_box[2] = 0; h->tune_steps = 0;
  h->dE_mode = 2;
  h->word_split = 1;
  h->wl = nullptr; h->comm = nullptr;
  h->word_epoch = 4;
#define BRW_CREATE_CUDA(x) do { if (brw_cuda_check((x), #x)) { brawl_cuda_destroy(h); return 1; } } while (0)
  BRW_CREATE_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  BRW_CREATE_CUDA(cudaMalloc(&h->d_lat, (size_t)g.n_sites * n_replicas));
  BRW_CREATE_CUDA(cudaMemsetAsync(h->d_lat, 0, (size_t)g.n_sites * n_replicas, h->stream));
  BRW_CREATE_CUDA(cudaMalloc(&h->d_V, sizeof(double) * S * S * n_shells));
  BRW_CREATE_CUDA(cudaMemcpyAsync(h->d_V, V, sizeof(double) * S * S * n_shells, cudaMemcpyHostToDevice, h->stream));
  BRW_CREATE_CUDA(cudaMalloc(&h->d_flag, sizeof(int)));
  BRW_CREATE_CUDA(cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
  BRW_CREATE_CUDA(cudaStreamSynchronize(h->stream));
  h->hV = new double[S * S * n_shells];
  memcpy(h->hV, V, sizeof(double) * S * S * n_shells);
  *out = h;
  return 0;
}

static void brw_free_plan(BrwPlan *pl) {
  if (!pl) return;
  cudaFree(pl->d_classes); cudaFree(pl->d_disp); cudaFree(pl->d_off); cudaFree(pl->d_Vrep);
  cudaFree(pl->d_att); cudaFree(pl->d_acc); cudaFree(pl->d_dE);
  delete pl;
}

static void brw_wl_free(brawl_cuda_ctx *h);
static void brw_comm_free(brawl_cuda_ctx *h);
extern "C" int brawl_cuda_destroy(brawl_cuda_t *h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  brw_wl_free(h); brw_comm_free(h);
  cudaFree(h->d_lat); cudaFree(h->d_V); cudaFree(h->d_stage); cudaFree(h->d_scratch); cudaFree(h->d_flag);
  cudaFree(h->d_beta); cudaFree(h->d_small); cudaFree(h->d_order);
  brw_free_plan((BrwPlan *)h->mc_plan[0]); brw_free_plan((BrwPlan *)h->mc_plan[1]);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete[] h->hV;
  delete h;
  return 0;
}
extern "C" int brawl_cuda_set_stream(brawl_cuda_t *h, void *s) {
  BRW_ENTER(h);
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  h->stream = s ? (cudaStream_t)s : h->own_stream;
  return 0;
}
extern "C" int brawl_cuda_synchronize(brawl_cuda_t *h) {
  BRW_ENTER(h);
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}
extern "C" int brawl_cuda_info(brawl_cuda_t *h, int64_t *n_atoms, int64_t *grid_bytes, int *z_total, int *n_replicas) {
  if (!h) return brw_fail("null handle");
  if (n_atoms) *n_atoms = h->g.n_sites;
  if (grid_bytes) *grid_bytes = h->grid_cells;
  if (z_total) *z_total = h->g.ztot;
  if (n_replicas) *n_replicas = h->n_replicas;
  return 0;
}

static int brw_check_flag(brawl_cuda_ctx *h, const char *what) {
  int f = 0;
  BRW_CUDA(cudaMemcpyAsync(&f, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  if (f) {
    BRW_CUDA(cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->stream));
    if (f & 1) return brw_fail("%s: species outside 1..n_species on a lattice site", what);
    if (f & 2) return brw_fail("%s: non-zero entry on a cell that is not a lattice site", what);
    return brw_fail("%s: index does not address a lattice site", what);
  }
  return 0;
}

// small device buffer for scalars / results
static int brw_small(brawl_cuda_ctx *h, size_t bytes) {
  if (h->small_bytes >= bytes) return 0;
  if (h->d_small) BRW_CUDA(cudaFree(h->d_small));
  h->d_small = nullptr; h->small_bytes = 0;
  size_t nb = bytes < 65536 ? 65536 : bytes;
  BRW_CUDA(cudaMalloc(&h->d_small, nb));
  h->small_bytes = nb;
  return 0;
}

// ---- configuration transfer -------------------------------------------------------------------------
extern "C" int brawl_cuda_set_config(brawl_cuda_t *h, int first, int n, const int8_t *grids) {
  BRW_ENTER(h);
  if (!grids) return brw_fail("null grid pointer");
  if (n < 1 || first < 0 || first + n > h->n_replicas) return brw_fail("replica range [%d,%d) out of [0,%d)", first, first + n, h->n_replicas);
  // chunk so that the staging buffer stays <= 256 MiB
  int per = (int)std::max<int64_t>(1, (256LL << 20) / h->grid_cells);
  for (int r0 = 0; r0 < n; r0 += per) {
    int m = std::min(per, n - r0);
    size_t bytes = (size_t)h->grid_cells * m;
    if (brw_ensure_stage(h, bytes)) return 1;
    BRW_CUDA(cudaMemcpyAsync(h->d_stage, grids + (size_t)r0 * h->grid_cells, bytes, cudaMemcpyHostToDevice, h->stream));
    if (h->g.gx % 16 == 0)
      brw_pack16_kernel<<<grid_for((long)(bytes / 16), 256), 256, 0, h->stream>>>(h->g, h->d_stage, h->d_lat + (size_t)(first + r0) * h->g.n_sites, m, h->d_flag);
    else
      brw_pack_kernel<<<grid_for((long)bytes, 256), 256, 0, h->stream>>>(h->g, h->d_stage, h->d_lat + (size_t)(first + r0) * h->g.n_sites, m, h->d_flag);
    BRW_LAUNCH_CHECK("brw_pack_kernel");
  }
  return brw_check_flag(h, "set_config");
}
extern "C" int brawl_cuda_get_config(brawl_cuda_t *h, int first, int n, int8_t *grids) {
  BRW_ENTER(h);
  if (!grids) return brw_fail("null grid pointer");
  if (n < 1 || first < 0 || first + n > h->n_replicas) return brw_fail("replica range [%d,%d) out of [0,%d)", first, first + n, h->n_replicas);
  int per = (int)std::max<int64_t>(1, (256LL << 20) / h->grid_cells);
  for (int r0 = 0; r0 < n; r0 += per) {
    int m = std::min(per, n - r0);
    size_t bytes = (size_t)h->grid_cells * m;
    if (brw_ensure_stage(h, bytes)) return 1;
    if (h->g.gx % 16 == 0)
      brw_unpack16_kernel<<<grid_for((long)(bytes / 16), 256), 256, 0, h->stream>>>(h->g, h->d_lat + (size_t)(first + r0) * h->g.n_sites, h->d_stage, m);
    else
      brw_unpack_kernel<<<grid_for((long)bytes, 256), 256, 0, h->stream>>>(h->g, h->d_lat + (size_t)(first + r0) * h->g.n_sites, h->d_stage, m);
    BRW_LAUNCH_CHECK("brw_unpack_kernel");
    BRW_CUDA(cudaMemcpyAsync(grids + (size_t)r0 * h->grid_cells, h->d_stage, bytes, cudaMemcpyDeviceToHost, h->stream));
    BRW_CUDA(cudaStreamSynchronize(h->stream));
  }
  return 0;
}
// Compact host buffers: [n][n_atoms] bytes, species 0..S-1 in the device's own site order (z, then compact y, then
// compact x -- the order brawl_cuda_lattice_ptr exposes).  A quarter (bcc) / half (fcc) of the bytes of the reference's
// `config` grid (src/shared_data.f90:30): the form a driver keeps between annealing segments or writes to a checkpoint
// when PCIe or host-memory bandwidth is shared by many GPUs.
__global__ void brw_validate_lattice_kernel(const uint8_t *__restrict__ lat, long n, int S, int *flag) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long n16 = n >> 4;
  bool bad = false;
  if (i < n16) {
    const uint4 v = reinterpret_cast<const uint4 *>(lat)[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
      for (int b = 0; b < 4; b++) bad |= ((w[k] >> (8 * b)) & 255u) >= (uint32_t)S;
  }
  if (i == 0) for (long j = n16 << 4; j < n; j++) bad |= lat[j] >= S;
  if (bad) atomicOr(flag, 1);
}
extern "C" int brawl_cuda_set_lattice(brawl_cuda_t *h, int first, int n, const uint8_t *sites) {
  BRW_ENTER(h);
  if (!sites) return brw_fail("null lattice pointer");
  if (n < 1 || first < 0 || first + n > h->n_replicas) return brw_fail("replica range [%d,%d) out of [0,%d)", first, first + n, h->n_replicas);
  const size_t bytes = (size_t)h->g.n_sites * n;
  uint8_t *dst = h->d_lat + (size_t)first * h->g.n_sites;
  BRW_CUDA(cudaMemcpyAsync(dst, sites, bytes, cudaMemcpyHostToDevice, h->stream));
  brw_validate_lattice_kernel<<<grid_for((long)(bytes / 16) + 1, 256), 256, 0, h->stream>>>(dst, (long)bytes, h->g.S, h->d_flag);
  BRW_LAUNCH_CHECK("brw_validate_lattice_kernel");
  return brw_check_flag(h, "set_lattice");
}
extern "C" int brawl_cuda_get_lattice(brawl_cuda_t *h, int first, int n, uint8_t *sites) {
  BRW_ENTER(h);
  if (!sites) return brw_fail("null lattice pointer");
  if (n < 1 || first < 0 || first + n > h->n_replicas) return brw_fail("replica range [%d,%d) out of [0,%d)", first, first + n, h->n_replicas);
  BRW_CUDA(cudaMemcpyAsync(sites, h->d_lat + (size_t)first * h->g.n_sites, (size_t)h->g.n_sites * n, cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}
// ---- start states on the device ---------------------------------------------------------------
// A uniformly random arrangement of the species multiset on the lattice sites -- the start state initial_setup
// (src/initialise.F90:434-617) and the WL re-randomisation (src/wang-landau.F90:671-674) produce -- for a batch of
// replicas at once: every site draws a 44-bit Philox key (counter = site, replica, offset), the (replica, key) pairs
// are radix-sorted (CUB, a library call: this is one-off set-up, not the hot path) carrying the species template
// 0..0 1..1 2..2 as values, and the sorted values ARE the compact lattices.  Key collisions (expected 0.5 pairs in
// 4 M sites) keep template order: a bias of order 1e-7 per site pair.
struct BrwCum { long long c[BRW_MAX_SPECIES + 1]; };
__global__ void __launch_bounds__(256) brw_shuffle_keys_kernel(long n_sites, int n_rep, int first, BrwCum cum, int S,
                                                               uint32_t k0, uint32_t k1, uint32_t off_lo, uint32_t off_hi,
                                                               unsigned long long *__restrict__ keys, uint8_t *__restrict__ vals) {
  const long total = n_sites * n_rep;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / n_sites, c = i - r * n_sites;
    const BrwPhilox4 p = brw_philox((uint32_t)c, (uint32_t)((unsigned long long)c >> 32) ^ 0x0A000000u, (uint32_t)(first + r),
                                    off_lo, k0, k1 ^ off_hi);
    keys[i] = ((unsigned long long)r << 44) | ((unsigned long long)p.x << 12) | (p.y >> 20);
    int sp = 0;
    while (sp < S - 1 && c >= cum.c[sp + 1]) sp++;
    vals[i] = (uint8_t)sp;
  }
}
extern "C" int brawl_cuda_random_config(brawl_cuda_t *h, int first, int n, const int64_t *species_count, uint64_t seed,
                                        uint64_t offset) {
  BRW_ENTER(h);
  if (!species_count) return brw_fail("null species_count");
  if (n < 1 || first < 0 || first + n > h->n_replicas) return brw_fail("replica range [%d,%d) out of [0,%d)", first, first + n, h->n_replicas);
  const BrwGeom &g = h->g;
  BrwCum cum;
  cum.c[0] = 0;
  for (int s = 0; s < g.S; s++) {
    if (species_count[s] < 0) return brw_fail("negative count for species %d", s + 1);
    cum.c[s + 1] = cum.c[s] + species_count[s];
  }
  if (cum.c[g.S] != (long long)g.n_sites)
    return brw_fail("species counts sum to %lld, the lattice has %lld sites", cum.c[g.S], (long long)g.n_sites);
  // chunks of <= 2^26 elements (and <= 2^19 replicas: the replica id sits above the 44 key bits)
  int per = (int)std::max<long long>(1, std::min<long long>((1LL << 26) / (long long)g.n_sites, 1LL << 19));
  for (int r0 = 0; r0 < n; r0 += per) {
    const int m = std::min(per, n - r0);
    const size_t ne = (size_t)g.n_sites * m;
    size_t tmp_bytes = 0;
    unsigned long long *kin = nullptr, *kout = nullptr;
    uint8_t *vin = nullptr, *vout = h->d_lat + (size_t)(first + r0) * g.n_sites;
    int rep_bits = 0;
    while ((1 << rep_bits) < m) rep_bits++;
    BRW_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kin, kout, vin, vout, (int)ne, 0, 44 + rep_bits, h->stream));
    const size_t tmp_al = (tmp_bytes + 255) & ~(size_t)255;
    if (brw_ensure_scratch(h, 2 * ne * sizeof(unsigned long long))) return 1;
    if (brw_ensure_stage(h, ne + 256 + tmp_al)) return 1;
    kin = (unsigned long long *)h->d_scratch;
    kout = kin + ne;
    vin = (uint8_t *)h->d_stage;
    void *tmp = (uint8_t *)h->d_stage + ((ne + 255) & ~(size_t)255);
    brw_shuffle_keys_kernel<<<grid_for((long)ne, 256), 256, 0, h->stream>>>((long)g.n_sites, m, first + r0, cum, g.S, (uint32_t)seed,
                                                                           (uint32_t)(seed >> 32), (uint32_t)offset,
                                                                           (uint32_t)(offset >> 32), kin, vin);
    BRW_LAUNCH_CHECK("brw_shuffle_keys_kernel");
    BRW_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kin, kout, vin, vout, (int)ne, 0, 44 + rep_bits, h->stream));
  }
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

// ---- atomic long-range order: site occupancies (store_state, src/analytics.f90:43-64) ---------------------------
__global__ void __launch_bounds__(256) brw_store_state_kernel(long n_sites, int S, int n_rep, const uint8_t *__restrict__ lat,
                                                              uint32_t *__restrict__ order) {
  const long total = n_sites * n_rep;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / n_sites, c = i - r * n_sites;
    order[(r * S + lat[i]) * n_sites + c] += 1u;          // one thread per (replica, site): no atomics needed
  }
}
// reference layout of `order`: (species, basis = 1, x, y, z), species fastest; off-site cells 0
__global__ void __launch_bounds__(256) brw_order_unpack_kernel(BrwGeom g, const uint32_t *__restrict__ order, double *__restrict__ out) {
  const long cells = (long)g.gx * g.gy * g.gz, total = cells * g.S;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long c = i / g.S;
    const int s = (int)(i - c * g.S);
    const int x = (int)(c % g.gx), y = (int)((c / g.gx) % g.gy), z = (int)(c / ((long)g.gx * g.gy));
    out[i] = brw_is_site(g, x, y, z) ? (double)order[(long)s * g.n_sites + brw_grid_to_compact(g, x, y, z)] : 0.0;
  }
}
extern "C" int brawl_cuda_store_state(brawl_cuda_t *h, int first, int n) {
  BRW_ENTER(h);
  if (n < 1 || first < 0 || first + n > h->n_replicas) return brw_fail("replica range [%d,%d) out of [0,%d)", first, first + n, h->n_replicas);
  const BrwGeom &g = h->g;
  const size_t per = (size_t)g.S * g.n_sites;
  if (!h->d_order) {
    BRW_CUDA(cudaMalloc(&h->d_order, per * h->n_replicas * sizeof(uint32_t)));
    BRW_CUDA(cudaMemsetAsync(h->d_order, 0, per * h->n_replicas * sizeof(uint32_t), h->stream));
  }
  brw_store_state_kernel<<<grid_for((long)g.n_sites * n, 256), 256, 0, h->stream>>>((long)g.n_sites, g.S, n, h->d_lat + (size_t)first * g.n_sites,
                                                                                  h->d_order + (size_t)first * per);
  BRW_LAUNCH_CHECK("brw_store_state_kernel");
  return 0;
}
extern "C" int brawl_cuda_get_order(brawl_cuda_t *h, int replica, double *order, int reset) {
  BRW_ENTER(h);
  BRW_REPLICA(h, replica);
  if (!order) return brw_fail("null order pointer");
  const BrwGeom &g = h->g;
  const size_t per = (size_t)g.S * g.n_sites, n_out = (size_t)h->grid_cells * g.S;
  if (!h->d_order) {                                        // nothing stored yet: all zeros
    memset(order, 0, n_out * sizeof(double));
    return 0;
  }
  if (brw_ensure_scratch(h, n_out * sizeof(double))) return 1;
  brw_order_unpack_kernel<<<grid_for((long)n_out, 256), 256, 0, h->stream>>>(g, h->d_order + (size_t)replica * per, h->d_scratch);
  BRW_LAUNCH_CHECK("brw_order_unpack_kernel");
  BRW_CUDA(cudaMemcpyAsync(order, h->d_scratch, n_out * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (reset) BRW_CUDA(cudaMemsetAsync(h->d_order + (size_t)replica * per, 0, per * sizeof(uint32_t), h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int brawl_cuda_copy_replica(brawl_cuda_t *h, int src, int dst) {
  BRW_ENTER(h);
  BRW_REPLICA(h, src); BRW_REPLICA(h, dst);
  if (src == dst) return 0;
  BRW_CUDA(cudaMemcpyAsync(h->d_lat + (size_t)dst * h->g.n_sites, h->d_lat + (size_t)src * h->g.n_sites, h->g.n_sites,
                           cudaMemcpyDeviceToDevice, h->stream));
  return 0;
}

// ---- Hamiltonian ---------------------------------------------------------------------------------------
// device-side: energies of replicas [first, first+n) into d_out[n] (no host sync)
static int brw_total_energy_dev(brawl_cuda_ctx *h, int first, int n, int exact, double *d_out) {
  const BrwGeom &g = h->g;
  if (exact) {
    int per = (int)std::max<long>(1, (long)((256LL << 20) / ((long)g.n_sites * 8)));
    for (int r0 = 0; r0 < n; r0 += per) {
      int m = std::min(per, n - r0);
      if (brw_ensure_scratch(h, (size_t)g.n_sites * m * sizeof(double))) return 1;
      const uint8_t *L = h->d_lat + (size_t)(first + r0) * g.n_sites;
      brw_site_energy_kernel<<<grid_for((long)g.n_sites * m, 256), 256, 0, h->stream>>>(g, h->d_V, L, h->d_scratch, m);
      BRW_LAUNCH_CHECK("brw_site_energy_kernel");
      brw_ordered_sum_kernel<<<m, 256, 0, h->stream>>>(h->d_scratch, g.n_sites, d_out + r0);
      BRW_LAUNCH_CHECK("brw_ordered_sum_kernel");
    }
  } else {
    // tiled kernel (compile-time gathers from shared memory) where it is instantiated and the lattice is a whole
    // number of 32 x 16 x 8 tiles; else one CTA per compact row, grid-strided
    typedef void (*BrwETileKernel)(BrwGeom, const double *, const uint8_t *, double *, int, int);
    BrwETileKernel tk = nullptr;
    if (g.S <= 5 && g.cx % 32 == 0 && g.cy % 16 == 0 && g.cz % 8 == 0 && !h->disable_fast) {
      if (g.lattice == 1 && g.n_shells == 4) tk = brw_energy_tile_kernel<1, 4>;
      else if (g.lattice == 1 && g.n_shells == 6) tk = brw_energy_tile_kernel<1, 6>;
      else if (g.lattice == 2 && g.n_shells == 4) tk = brw_energy_tile_kernel<2, 4>;
      else if (g.lattice == 2 && g.n_shells == 6) tk = brw_energy_tile_kernel<2, 6>;
    }
    if (tk) {
      const int ntx = g.cx / 32, nty = g.cy / 16, ntz = g.cz / 8, ntile = ntx * nty * ntz;
      if (brw_ensure_scratch(h, (size_t)ntile * n * sizeof(double))) return 1;
      tk<<<dim3(ntile, n), 256, 0, h->stream>>>(g, h->d_V, h->d_lat + (size_t)first * g.n_sites, h->d_scratch, ntx, nty);
      BRW_LAUNCH_CHECK("brw_energy_tile_kernel");
      brw_tree_final_kernel<<<n, 128, 0, h->stream>>>(h->d_scratch, ntile, d_out, n);
      BRW_LAUNCH_CHECK("brw_tree_final_kernel");
      return 0;
    }
    int n_rows = g.cy * g.cz;
    int nblk = std::max(1, std::min(n_rows, n >= 148 ? 8 : (n >= 8 ? 148 : 1184)));
    if (brw_ensure_scratch(h, (size_t)nblk * n * sizeof(double))) return 1;
    dim3 grid(nblk, n);
    int threads = g.cx >= 128 ? 128 : (g.cx > 32 ? 64 : 32);
    brw_energy_partial_kernel<<<grid, threads, sizeof(double) * g.S * g.S * g.n_shells, h->stream>>>(
        g, h->d_V, h->d_lat + (size_t)first * g.n_sites, h->d_scratch, nblk);
    BRW_LAUNCH_CHECK("brw_energy_partial_kernel");
    brw_tree_final_kernel<<<n, 128, 0, h->stream>>>(h->d_scratch, nblk, d_out, n);
    BRW_LAUNCH_CHECK("brw_tree_final_kernel");
  }
  return 0;
}
extern "C" int brawl_cuda_total_energy(brawl_cuda_t *h, int first, int n, int exact, double *energies) {
  BRW_ENTER(h);
  if (!energies) return brw_fail("null output pointer");
  if (n < 1 || first < 0 || first + n > h->n_replicas) return brw_fail("replica range [%d,%d) out of [0,%d)", first, first + n, h->n_replicas);
  if (brw_small(h, sizeof(double) * n)) return 1;
  if (brw_total_energy_dev(h, first, n, exact, (double *)h->d_small)) return 1;
  BRW_CUDA(cudaMemcpyAsync(energies, h->d_small, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}
extern "C" int brawl_cuda_site_energies(brawl_cuda_t *h, int replica, double *out) {
  BRW_ENTER(h);
  BRW_REPLICA(h, replica);
  if (!out) return brw_fail("null output pointer");
  const BrwGeom &g = h->g;
  size_t nb = ((size_t)g.n_sites + (size_t)h->grid_cells) * sizeof(double);
  if (brw_ensure_scratch(h, nb)) return 1;
  double *d_e = h->d_scratch, *d_grid = h->d_scratch + g.n_sites;
  brw_site_energy_kernel<<<grid_for(g.n_sites, 256), 256, 0, h->stream>>>(g, h->d_V, h->d_lat + (size_t)replica * g.n_sites, d_e, 1);
  BRW_LAUNCH_CHECK("brw_site_energy_kernel");
  brw_site_energy_to_grid_kernel<<<grid_for(h->grid_cells, 256), 256, 0, h->stream>>>(g, d_e, d_grid);
  BRW_LAUNCH_CHECK("brw_site_energy_to_grid_kernel");
  BRW_CUDA(cudaMemcpyAsync(out, d_grid, (size_t)h->grid_cells * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}
// setup%nbr_energy(config, 1, i, j, k) for one site, optionally with another species on it (the operator takes the species
// of the centre as an argument, src/bw_hamiltonian.f90:898-1262): species = 0 uses the occupant
__global__ void brw_one_site_energy_kernel(BrwGeom g, const double *__restrict__ V, const uint8_t *__restrict__ lat, int x, int y,
                                           int z, int species, double *out) {
  const int c = brw_grid_to_compact(g, x, y, z);
  *out = brw_site_energy(g, V, x, y, z, species > 0 ? species - 1 : lat[c], BrwPlainSpec{lat});
}
extern "C" int brawl_cuda_nbr_energy(brawl_cuda_t *h, int replica, int x, int y, int z, int species, double *energy) {
  BRW_ENTER(h);
  BRW_REPLICA(h, replica);
  if (!energy) return brw_fail("null output pointer");
  const BrwGeom &g = h->g;
  if (x < 0 || y < 0 || z < 0 || x >= g.gx || y >= g.gy || z >= g.gz) return brw_fail("site (%d,%d,%d) outside the %dx%dx%d grid", x, y, z, g.gx, g.gy, g.gz);
  const bool on = g.lattice == 0 ? true : g.lattice == 1 ? ((x & 1) == (z & 1) && (y & 1) == (z & 1)) : (((x + y + z) & 1) == 0);
  if (!on) return brw_fail("(%d,%d,%d) is not a lattice site", x, y, z);
  if (species < 0 || species > g.S) return brw_fail("species %d out of 0..%d", species, g.S);
  if (brw_small(h, sizeof(double))) return 1;
  brw_one_site_energy_kernel<<<1, 1, 0, h->stream>>>(g, h->d_V, h->d_lat + (size_t)replica * g.n_sites, x, y, z, species, (double *)h->d_small);
  BRW_LAUNCH_CHECK("brw_one_site_energy_kernel");
  BRW_CUDA(cudaMemcpyAsync(energy, h->d_small, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}
extern "C" int brawl_cuda_pair_dE(brawl_cuda_t *h, int replica, int64_t n, const int32_t *i1, const int32_t *i2, double *dE) {
  BRW_ENTER(h);
  BRW_REPLICA(h, replica);
  if (n < 0 || (n > 0 && (!i1 || !i2 || !dE))) return brw_fail("bad pair arrays");
  if (n == 0) return 0;
  size_t nb = (size_t)n * (2 * sizeof(int32_t) + sizeof(double));
  if (brw_ensure_scratch(h, nb)) return 1;
  double *d_dE = h->d_scratch;
  int32_t *d_i1 = (int32_t *)(d_dE + n), *d_i2 = d_i1 + n;
  BRW_CUDA(cudaMemcpyAsync(d_i1, i1, n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  BRW_CUDA(cudaMemcpyAsync(d_i2, i2, n * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  brw_pair_dE_kernel<<<grid_for(n, 128), 128, 0, h->stream>>>(h->g, h->d_V, h->d_lat + (size_t)replica * h->g.n_sites, n, d_i1, d_i2, d_dE, h->d_flag);
  BRW_LAUNCH_CHECK("brw_pair_dE_kernel");
  BRW_CUDA(cudaMemcpyAsync(dE, d_dE, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  return brw_check_flag(h, "pair_dE");
}

// ---- Metropolis replay -----------------------------------------------------------------------------------
extern "C" int brawl_cuda_metropolis_replay_sampled(brawl_cuda_t *h, int replica, double beta, int64_t n_trials,
                                                    int64_t n_sample, int nbr_swap, uint32_t *mt, int64_t *n_accept,
                                                    double *energies) {
  BRW_ENTER(h);
  BRW_REPLICA(h, replica);
  if (!mt) return brw_fail("null MT19937 state");
  if (n_trials < 0 || n_sample < 0) return brw_fail("negative trial count");
  if (n_sample > 0 && !energies) return brw_fail("null energies output");
  if (mt[624] > 625) return brw_fail("corrupt MT19937 state (mti=%u)", mt[624]);
  const BrwGeom &g = h->g;
  int64_t ns = n_sample > 0 ? n_trials / n_sample : 0;
  size_t nb = 625 * sizeof(uint32_t) + 16 + (size_t)ns * sizeof(double);
  if (brw_small(h, nb)) return 1;
  if (brw_ensure_scratch(h, (size_t)g.n_sites * sizeof(double))) return 1;
  unsigned long long *d_acc = (unsigned long long *)h->d_small;
  double *d_en = (double *)(d_acc + 1);
  uint32_t *d_mt = (uint32_t *)(d_en + ns);
  BRW_CUDA(cudaMemcpyAsync(d_mt, mt, 625 * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  brw_metropolis_replay_kernel<<<1, 256, 0, h->stream>>>(g, h->d_V, h->d_lat + (size_t)replica * g.n_sites, beta, (long)n_trials,
                                                         (long)n_sample, nbr_swap, d_mt, d_acc, d_en, h->d_scratch);
  BRW_LAUNCH_CHECK("brw_metropolis_replay_kernel");
  unsigned long long acc = 0;
  BRW_CUDA(cudaMemcpyAsync(mt, d_mt, 625 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaMemcpyAsync(&acc, d_acc, sizeof acc, cudaMemcpyDeviceToHost, h->stream));
  if (ns > 0) BRW_CUDA(cudaMemcpyAsync(energies, d_en, ns * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  if (n_accept) *n_accept = (int64_t)acc;
  return 0;
}
extern "C" int brawl_cuda_metropolis_replay(brawl_cuda_t *h, int replica, double beta, int64_t n_trials, int nbr_swap,
                                            uint32_t *mt, int64_t *n_accept) {
  return brawl_cuda_metropolis_replay_sampled(h, replica, beta, n_trials, 0, nbr_swap, mt, n_accept, nullptr);
}

// ---- production Metropolis ---------------------------------------------------------------------------------
typedef void (*BrwFastKernel)(BrwGeom, BrwBoxParams, uint8_t *, const double *, const double *, const int4 *,
                              const int4 *, uint32_t, uint32_t, uint32_t, int, unsigned long long *, unsigned long long *,
                              double *);
struct BrwFastEntry { int lat, nsh, px, py, maxt; BrwFastKernel fn, fn_screen; };
// byte-lattice epoch kernels of the same geometries (epoch_byte_metropolis.cuh), instantiated in byte_epoch_kernels.cu
void *brw_byte_epoch_kernel_lookup(int lat, int nsh, int px, int py, int maxt, int epoch_k, int exact);
int brw_byte_epoch_pitch(int lat, int nsh);
// MAXT = launch bound: CTAs of <= 384 threads (e.g. the 128^3 single chain, 352 threads) may use up to
// 168 registers/thread, CTAs of <= 768 threads (e.g. one 32^3-cell replica per CTA, 736 threads) 80.
#define BRW_FAST(LAT, NSH, PX, PY, MAXT) {LAT, NSH, PX, PY, MAXT, brw_box_metropolis_fast_kernel<LAT, NSH, PX, PY, false, MAXT>, \
                                          brw_box_metropolis_fast_kernel<LAT, NSH, PX, PY, true, MAXT>}
static const BrwFastEntry brw_fast_table[] = {
    BRW_FAST(1, 4, 32, 32, 512), BRW_FAST(1, 4, 32, 32, 768), BRW_FAST(1, 6, 32, 32, 512), BRW_FAST(1, 6, 32, 32, 768),
    BRW_FAST(2, 4, 32, 64, 512), BRW_FAST(2, 4, 32, 64, 768), BRW_FAST(2, 6, 32, 64, 512), BRW_FAST(2, 6, 32, 64, 768),
};

// Word-lattice kernels with the dense decomposition (word_metropolis.cuh): fixed box and margin per entry.
#ifndef BRW_TMA_LOAD
#define BRW_TMA_LOAD 1    // epoch kernel: box load through cp.async.bulk (A/B switch)
#endif
#ifndef BRW_BYTE_PDL
#define BRW_BYTE_PDL 1    // programmatic dependent launch for the byte-lattice epoch kernels (A/B switch)
#endif
#ifndef BRW_NLIMB
#define BRW_NLIMB 4         // signed 8-bit digits of the fixed-point table (word_metropolis.cuh)
#endif
#ifndef BRW_PAIRW
#define BRW_PAIRW true      // pair words (two sites per LDS.32); false: one-hot byte word per site
#endif
// pitchz < bzc: consecutive z layers share their frozen margin planes (box extent bzc, layer pitch pitchz = bzc - margin);
// split: two independent warp groups per CTA on disjoint z zones (word_metropolis.cuh, SPLIT)
// epoch > 0: counts cached over epochs of that many steps (epoch_metropolis.cuh); pitchxc < bxc: x margins shared as well
struct BrwWordEntry { int lat, nsh, bxc, byc, bzc, margin, pxp, plp, nlimb; BrwFastKernel fn, fn_exact; int pitchz; bool split; int epoch = 0; int pitchxc = 0; };
static const BrwWordEntry brw_word_table[] = {
    // bcc, 4 shells, box 64x64x32 (doubled-grid units): 32 warps x 28 = 896 trials per step
    {1, 4, 32, 32, 32, 4, 32, 1024, BRW_NLIMB, brw_box_metropolis_word_kernel<1, 4, 32, 32, 32, 4, 32, 1024, BRW_NLIMB, false, BRW_PAIRW>,
     brw_box_metropolis_word_kernel<1, 4, 32, 32, 32, 4, 32, 1024, BRW_NLIMB, true, BRW_PAIRW>, 32, false},
    // box 64x64x28: 30 warps x 28 = 840 trials per step; 9 z-layers of a 256-plane lattice give 144 boxes (of 148 SMs)
    {1, 4, 32, 32, 28, 4, 32, 1024, BRW_NLIMB, brw_box_metropolis_word_kernel<1, 4, 32, 32, 28, 4, 32, 1024, BRW_NLIMB, false, BRW_PAIRW>,
     brw_box_metropolis_word_kernel<1, 4, 32, 32, 28, 4, 32, 1024, BRW_NLIMB, true, BRW_PAIRW>, 28, false},
    // box 64x64x32 at z pitch 28 (shared margin planes), two warp groups of 12 + 18 warps on z planes {0,1} / {3,4,5}:
    // 840 trials per step like the 64x64x28 box, but the groups' gather and arithmetic phases overlap
    {1, 4, 32, 32, 32, 4, 32, 1024, BRW_NLIMB, brw_box_metropolis_word_kernel<1, 4, 32, 32, 32, 4, 32, 1024, BRW_NLIMB, false, BRW_PAIRW, true>,
     brw_box_metropolis_word_kernel<1, 4, 32, 32, 32, 4, 32, 1024, BRW_NLIMB, true, BRW_PAIRW, true>, 28, true},
    // box 68x64x32 at a pitch of 64x64x28 (margins shared in x and z: 15 sites per class row), 32 warps x 30 lanes = 960
    // trials per step, neighbour counts cached over epochs of 8 / 4 steps
    {1, 4, 34, 32, 32, 4, 34, 1088, BRW_NLIMB, brw_box_metropolis_epoch_kernel<4, 34, 32, 32, 4, 34, 1088, BRW_NLIMB, false, 8>,
     brw_box_metropolis_epoch_kernel<4, 34, 32, 32, 4, 34, 1088, BRW_NLIMB, true, 8>, 28, false, 8, 32},
    {1, 4, 34, 32, 32, 4, 34, 1088, BRW_NLIMB, brw_box_metropolis_epoch_kernel<4, 34, 32, 32, 4, 34, 1088, BRW_NLIMB, false, 4>,
     brw_box_metropolis_epoch_kernel<4, 34, 32, 32, 4, 34, 1088, BRW_NLIMB, true, 4>, 28, false, 4, 32},
    {1, 4, 34, 32, 32, 4, 34, 1088, BRW_NLIMB, brw_box_metropolis_epoch_kernel<4, 34, 32, 32, 4, 34, 1088, BRW_NLIMB, false, 2>,
     brw_box_metropolis_epoch_kernel<4, 34, 32, 32, 4, 34, 1088, BRW_NLIMB, true, 2>, 28, false, 2, 32},
};

// launch helper for the warp-per-walker kernels: layout + opt-in shared memory
template <class K>
static int brw_walker_prepare(const BrwGeom &g, int extra_doubles, K kernel, BrwWalkerLayout &lay) {
  lay = brw_walker_layout(g, extra_doubles);
  if (lay.total() + sizeof(BrwWarpScratch) * BRW_WALKER_WARPS > 220 * 1024) {
    // too much per-walker state for shared memory: drop the table, then the staging
    lay.use_tab = 0; lay.tab_bytes = 0; lay.ksh_bytes = 0; lay.fast_delta = 0;
    if (lay.total() + sizeof(BrwWarpScratch) * BRW_WALKER_WARPS > 220 * 1024) return brw_fail("walker state does not fit in shared memory");
  }
  if (brw_cuda_check(cudaFuncSetAttribute((const void *)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total()), "cudaFuncSetAttribute")) return 1;
  return 0;
}


// Site-energy table of the epoch kernels (epoch_metropolis.cuh, epoch_byte_metropolis.cuh): X_n[a][s] = V_n(a, s) - V_n(a, 4) -
// (V_n(4, s) - V_n(4, 4)) for the species a of a cache row (S = 5; V_n(a, s) - V_n(3, s) for S = 4, plain V_n(a, s) below),
// rounded to fixed-point units (fix_scale) and split into three signed 8-bit digits.  |h - real| <= Z/2 units per cached value, four values per
// dE  =>  guard = 2 Z units (+ the f64 rounding slack against the reference association).  S <= 4: all four counts are
// explicit, energies relative to species 3 (its row is identically zero and not computed).
static void brw_epoch_tables(const brawl_cuda_ctx *h, BrwBoxParams &p, double umax) {
  const BrwGeom &g = h->g;
  const int S = g.S, NSH = g.n_shells;
  auto Vn = [&](int n, int centre, int nbr) { return h->hV[(n * S + nbr) * S + centre]; };
  p.h_rows = S == 5 ? 4 : 3;
  auto X = [&](int n, int a, int s2) {
    double x = Vn(n, a, s2);
    if (S == 5) x -= Vn(n, a, 4) + Vn(n, 4, s2) - Vn(n, 4, 4);
    else if (S == 4) x -= Vn(n, 3, s2);
    return x;
  };
  double xmax = 0.0, vmax = 0.0;
  for (int n = 0; n < NSH; n++) for (int a = 0; a < std::min(S, 4); a++) for (int s2 = 0; s2 < std::min(S, 4); s2++)
    xmax = std::max(xmax, std::fabs(X(n, a, s2)));
  for (int i = 0; i < S * S * NSH; i++) vmax = std::max(vmax, std::fabs(h->hV[i]));
  if (umax <= 0.0) umax = 2.0 * vmax;
  // unit of the fixed-point table: the largest entry maps to 8.3e6 (three signed 8-bit digits reach 127 * 65793 = 8.36e6),
  // or less where four cached values of up to Z * that could overflow the int32 dE
  const double top = std::min(8.3e6, 0.99 * 2147483647.0 / (4.0 * g.ztot));
  const double inv_unit = xmax > 0.0 ? top / xmax : 1.0;
  p.fix_scale = 1.0 / inv_unit;
  p.guard = 2.0 * g.ztot * p.fix_scale + 1e-9 * g.ztot * umax;
  p.guard2 = 1e-9 * g.ztot * vmax;                   // f64 count-based dE vs the reference association: rounding ~1e-16 * Z * max|V|
  p.gfix = (int)std::ceil(p.guard / p.fix_scale) + 1;
  std::memset(p.xdig, 0, sizeof p.xdig);
  for (int c = 0; c < 4; c++) {
    const int a = c;                                  // cache row <-> species c
    if (a >= S || c >= p.h_rows) continue;
    for (int n = 0; n < NSH; n++)
      for (int s2 = 0; s2 < std::min(S, 4); s2++) {
        long long v = std::llrint(X(n, a, s2) * inv_unit);
        for (int k = 0; k < 3; k++) {
          long long dgt = k == 2 ? v : ((v + 128) & 255) - 128;
          v = (v - dgt) >> 8;
          p.xdig[(n * 3 + k) * 4 + c] |= (int)((uint32_t)(uint8_t)(int8_t)dgt << (8 * s2));
        }
      }
  }
}

static int brw_build_plan(brawl_cuda_ctx *h, int nbr_swap, BrwPlan **out) {
  BrwPlan *&slot = *(BrwPlan **)&h->mc_plan[nbr_swap ? 1 : 0];
  if (slot && slot->valid) { *out = slot; return 0; }
  if (slot) { brw_free_plan(slot); slot = nullptr; }
  const BrwGeom &g = h->g;
  BrwPlan *pl = new BrwPlan();
  pl->nbr_swap = nbr_swap;
  BrwBoxParams &p = pl->p;
  // reach of the Hamiltonian
  int rmax = 0;
  for (int k = 0; k < g.ztot; k++) for (int c = 0; c < 3; c++) rmax = std::max(rmax, std::abs((int)g.off[k][c]));
  int m = rmax + (nbr_swap ? 1 : 0);
  m += m & 1;
  std::vector<std::array<int, 3>> first;
  {
    static const signed char sc[6][3] = {{0,0,1},{0,1,0},{1,0,0},{0,0,-1},{0,-1,0},{-1,0,0}};
    static const signed char bcc[8][3] = {{1,1,1},{1,1,-1},{1,-1,1},{1,-1,-1},{-1,1,1},{-1,1,-1},{-1,-1,1},{-1,-1,-1}};
    static const signed char fcc[12][3] = {{0,1,1},{0,1,-1},{0,-1,1},{0,-1,-1},{1,1,0},{1,-1,0},{1,0,1},{1,0,-1},{-1,1,0},{-1,-1,0},{-1,0,1},{-1,0,-1}};
    int n = g.lattice == 0 ? 6 : g.lattice == 1 ? 8 : 12;
    for (int i = 0; i < n; i++) {
      const signed char *e = g.lattice == 0 ? sc[i] : g.lattice == 1 ? bcc[i] : fcc[i];
      first.push_back({e[0], e[1], e[2]});
    }
  }
  // candidate period sets (smallest volume first); the box is sized with the first one, then every
  // candidate is scored for that box
  std::vector<std::vector<BrwModeChoice>> cands;
  if (g.lattice != 0)   // simple cubic wraps neighbours with modulus n (reference quirk): chain kernel only
    brw_candidate_modes(g, nbr_swap, first, 16, h->cubic_period_only != 0, 6, cands);
  bool feasible = !cands.empty();
  std::vector<BrwModeChoice> modes;
  if (feasible) modes = cands[0];
  int Pmax = 0;                       // largest period entry over the candidates: minimum box extent
  for (auto &cs : cands) for (auto &mc : cs) for (int d = 0; d < 3; d++) Pmax = std::max(Pmax, mc.P[d]);
  const int P = Pmax;
  int B[3] = {g.gx, g.gy, g.gz};
  if (feasible) {
    for (int d = 0; d < 3; d++) if (B[d] < 2 * m + P) feasible = false;
  }
  // box extents: user override or automatic halving of the longest edge, for a given shared-memory
  // footprint per site.  Returns 0 ok, 1 infeasible, 2 error (message set).
  const size_t budget = 200 * 1024;
  auto size_box = [&](double bytes_per_site, size_t fixed, int minP, const int *P0, int *Bo, const int *stop_at = nullptr) -> int {
    for (int d = 0; d < 3; d++) Bo[d] = d == 0 ? g.gx : d == 1 ? g.gy : g.gz;
    auto bytes = [&](const int *b) { return (size_t)((double)b[0] * b[1] * b[2] / (g.lattice == 1 ? 4 : 2) * bytes_per_site); };
    auto trials = [&](const int *b) { long M = 1; for (int d = 0; d < 3; d++) M *= (b[d] - 2 * m) / P0[d]; return M; };
    if (h->tune_box[0] > 0) {
      for (int d = 0; d < 3; d++) {
        int G = d == 0 ? g.gx : d == 1 ? g.gy : g.gz;
        Bo[d] = h->tune_box[d];
        if (Bo[d] < 2 * m + minP || (Bo[d] & 1) || G % Bo[d]) {
          brw_fail("box extent %d invalid for axis %d (grid %d, need even divisor >= %d)", Bo[d], d, G, 2 * m + minP);
          return 2;
        }
      }
      if (bytes(Bo) + fixed > budget) { brw_fail("box does not fit in shared memory"); return 2; }
      return 0;
    }
    for (;;) {
      long nbox = (long)(g.gx / Bo[0]) * (g.gy / Bo[1]) * (g.gz / Bo[2]) * h->n_replicas;
      bool too_big = bytes(Bo) + fixed > budget;
      bool want_more = nbox < 120 && trials(Bo) >= 512;
      if (!too_big && !want_more) return 0;
      if (!too_big && stop_at && Bo[0] == stop_at[0] && Bo[1] == stop_at[1] && Bo[2] == stop_at[2]) return 0;
      // halve the longest edge that can still be halved
      // (a box that has a specialised kernel is not halved out of it just to get more CTAs)
      auto has_fast = [&](const int *b) {
        for (const BrwFastEntry &fe : brw_fast_table)
          if (fe.lat == g.lattice && fe.nsh == g.n_shells && fe.px == (b[0] >> g.xs) && fe.py == (b[1] >> g.ys)) return true;
        return false;
      };
      int best = -1;
      for (int d = 0; d < 3; d++) {
        int nbd = Bo[d] / 2;
        if ((Bo[d] % 4) != 0 || nbd < 2 * m + minP) continue;
        if (!too_big && has_fast(Bo)) {
          int Bh[3] = {Bo[0], Bo[1], Bo[2]};
          Bh[d] = nbd;
          if (!has_fast(Bh)) continue;
        }
        if (best < 0 || Bo[d] >= Bo[best]) best = d;   // ties: z, then y, then x
      }
      if (best < 0) return too_big ? 1 : 0;
      Bo[best] /= 2;
    }
  };
  // word-lattice kernels with the dense decomposition: used when the box sized for 4-byte sites is one an entry is
  // instantiated for (dE_mode 0 -> its EXACT instantiation, 2 -> screened; dE_mode 1 and byte_layout keep the
  // byte-lattice path)
  const BrwWordEntry *we = nullptr;
  if (feasible && !nbr_swap && !h->cubic_period_only && !h->disable_fast && !h->byte_layout && h->dE_mode != 1 && g.S <= 5) {
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device);
    double best_rate = 0.0;
    for (const BrwWordEntry &e : brw_word_table) {
      if (e.lat != g.lattice || e.nsh != g.n_shells || e.margin < rmax + 1) continue;
      // S_o = o + {(0,0,0),(2,2,2)} + 4Z^3 must be an independent set of the interaction graph
      bool independent = true;
      for (int k = 0; k < g.ztot; k++) {
        const int a = brw_posmod(g.off[k][0], 4), b = brw_posmod(g.off[k][1], 4), c = brw_posmod(g.off[k][2], 4);
        if ((a == 0 && b == 0 && c == 0) || (a == 2 && b == 2 && c == 2)) independent = false;
      }
      if (!independent) continue;
      const int Be[3] = {(e.pitchxc ? e.pitchxc : e.bxc) << g.xs, e.byc << g.ys, e.bzc};    // x, y: box pitch
      if (e.epoch != h->word_epoch) continue;
      if (h->tune_box[0] > 0) {          // user box: must be exactly this entry's box (tune_box x, z = pitch)
        if (h->tune_box[0] != Be[0] || h->tune_box[1] != Be[1] || h->tune_box[2] != e.pitchz) continue;
        if (!e.epoch && e.split != (h->word_split != 0)) continue;
      } else if (e.split && h->word_split == 0) continue;
      // boxes tile x and y; along z they need not (the slab left over is visited in later phases: the origin is random)
      if (g.gx % Be[0] || g.gy % Be[1] || Be[2] > g.gz) continue;
      const int nbz = (g.gz - (e.bzc - e.pitchz)) / e.pitchz;
      const long n_boxes = (long)(g.gx / Be[0]) * (g.gy / Be[1]) * nbz * h->n_replicas;
      const int planes = (Be[2] - 2 * e.margin) / 4 - (e.split ? 1 : 0);
      const int rows = std::min(32, ((((Be[1] - 3 * e.margin) / 2) & ~3) / 4) * planes);
      const long trials_box = (long)rows * 2 * (((e.bxc << g.xs) - 2 * e.margin) / 4);
      // one CTA per SM; a step costs ~rows warp-gathers on the shared-memory pipe (overlapped with the arithmetic of
      // the other group's in the split kernels: measured +5 %)
      const long waves = (n_boxes + n_sm - 1) / n_sm;
      double rate = (double)n_boxes * trials_box / ((double)waves * rows) * (e.split ? 1.05 : 1.0);
      if (nbz * e.pitchz + (e.bzc - e.pitchz) == g.gz) rate *= 1.02;                       // prefer exact tilings when (nearly) equal
      if (rate > best_rate) { best_rate = rate; we = &e; }
    }
    if (we) {
      B[0] = (we->pitchxc ? we->pitchxc : we->bxc) << g.xs; B[1] = we->byc << g.ys; B[2] = we->pitchz;     // B = box pitch (origin stride)
      m = we->margin;
    }
  }
  if (feasible && !we) {
    size_t fixed = (size_t)g.S * g.S * g.n_shells * 16 * 8 + 2 * g.ztot * 4 + 2 * sizeof(BrwStepParams) + 32 * 8 + 64;
    const int st = size_box(1.0, fixed, P, modes[0].P, B);
    if (st == 2) { delete pl; return 1; }
    if (st == 1) feasible = false;
  }
  if (feasible) {
    p.m = m;
    {
      double vmax = 0.0;
      for (int i = 0; i < g.S * g.S * g.n_shells; i++) vmax = std::max(vmax, std::fabs(h->hV[i]));
      p.guard = 1e-9 * g.ztot * vmax;
    }
    for (int d = 0; d < 3; d++) {
      int G = d == 0 ? g.gx : d == 1 ? g.gy : g.gz;
      p.B[d] = B[d]; p.nb[d] = G / B[d];
    }
    if (we) p.nb[2] = (g.gz - (we->bzc - we->pitchz)) / we->pitchz;
    p.bxc = B[0] >> g.xs; p.byc = B[1] >> g.ys; p.bzc = we ? we->bzc : B[2];
    p.box_sites = p.bxc * p.byc * p.bzc;
    // is a specialised kernel instantiated for this lattice / shell count / box pitch?  Its CTAs hold at
    // most 768 threads, one trial each.
    bool has_fast = false;
    if (!nbr_swap && !h->disable_fast)
      for (const BrwFastEntry &fe : brw_fast_table)
        if (fe.lat == g.lattice && fe.nsh == g.n_shells && fe.px == p.bxc && fe.py == p.byc) has_fast = true;
    const int cap = has_fast ? 768 : 1 << 30;
    // score every candidate for this box: simultaneous trials per step, limited by the CTA size
    long best_score = -1;
    for (auto &cs : cands) {
      long worst = 1L << 40;
      for (auto &mc : cs) {
        long M = 1;
        for (int d = 0; d < 3; d++) M *= (B[d] - 2 * m) / mc.P[d];
        worst = std::min(worst, std::min<long>(M, cap));
      }
      if (worst > best_score) { best_score = worst; modes = cs; }
    }
    p.n_modes = (int)modes.size();
    std::vector<int4> classes, disp;
    int Mmax = 0, Mmin = 1 << 30;
    for (int q = 0; q < p.n_modes; q++) {
      BrwBoxMode &md = p.mode[q];
      for (int d = 0; d < 3; d++) { md.P[d] = modes[q].P[d]; md.A[d] = (B[d] - 2 * m) / md.P[d]; }
      // shrink the active region along its longest axis until the trials of one step fit the CTA
      while ((long)md.A[0] * md.A[1] * md.A[2] > cap) {
        int big = 0;
        for (int d = 1; d < 3; d++) if (md.A[d] > md.A[big]) big = d;
        md.A[big]--;
      }
      md.M = md.A[0] * md.A[1] * md.A[2];
      md.cls0 = (int)classes.size(); md.n_classes = (int)modes[q].classes.size();
      md.d0 = (int)disp.size(); md.n_disp = (int)modes[q].disp.size();
      classes.insert(classes.end(), modes[q].classes.begin(), modes[q].classes.end());
      disp.insert(disp.end(), modes[q].disp.begin(), modes[q].disp.end());
      Mmax = std::max(Mmax, md.M); Mmin = std::min(Mmin, md.M);
    }
    if (we) {
      // dense decomposition (word_metropolis.cuh, BrwDenseGeom): period 4, one warp per (A row, B row) pair
      p.n_modes = 1;
      BrwBoxMode &md = p.mode[0];
      md.P[0] = md.P[1] = md.P[2] = 4;
      md.A[0] = ((we->bxc << g.xs) - 2 * m) / 4;                  // sites per x-row of one sub-class
      md.A[1] = (((B[1] - 3 * m) / 2) & ~3) / 4;                  // rows per plane in one y half
      md.A[2] = (we->bzc - 2 * m) / 4;                            // planes per sub-class
      md.M = std::min(32, md.A[1] * (md.A[2] - (we->split ? 1 : 0))) * 2 * md.A[0];
      md.cls0 = md.d0 = 0; md.n_classes = 16; md.n_disp = 256;    // 16 residue classes mod 4, every (o, o') pair
      Mmax = md.M;
    }
    pl->Mmax = Mmax;
    p.boxes_per_replica = p.nb[0] * p.nb[1] * p.nb[2];
    p.v_entries = g.S * g.S * g.n_shells;
    // default: about four sweeps of the box per phase (amortises the box load/store); six for the word kernels,
    // whose steps are short enough for the box copies to be 9 % of a four-sweep phase
    int steps = h->tune_steps > 0 ? h->tune_steps : ((we ? 6 : 4) * p.box_sites + Mmax - 1) / Mmax;
    p.steps = std::max(8, std::min(steps, 512));
    if (we && we->epoch) p.steps = ((p.steps + we->epoch - 1) / we->epoch) * we->epoch;      // whole epochs
    // byte-lattice epoch kernels (epoch_byte_metropolis.cuh): every geometry with a specialised kernel but no dense-set
    // word kernel (fcc, 6-shell bcc), unless the caller asked for the one-gather-per-step kernels (byte_layout, dE_mode 1,
    // set_layout 2/3)
    const int byte_epoch = (!we && !nbr_swap && !h->disable_fast && !h->byte_layout && h->dE_mode != 1 && g.S <= 5 &&
                            h->word_epoch > 0 && Mmax <= 768) ? (h->word_epoch == 8 ? 8 : 4) : 0;
    if (byte_epoch && has_fast) {
      // a trial costs ~10x less than a gather-per-step trial: six sweeps per phase amortise the box copies
      if (h->tune_steps <= 0) steps = (6 * p.box_sites + Mmax - 1) / Mmax;
      p.steps = std::max(8, std::min(steps, 1024));
      p.steps = ((p.steps + byte_epoch - 1) / byte_epoch) * byte_epoch;
    }
    p.steps_a = p.steps;
    pl->M_a = 0;
    if (we && we->split) {
      // group A holds NPA of the NP - 1 working planes (12 of 30 warps for the 64x64x32 box) and was measured to idle
      // 5.3 % of the phase waiting for group B: give it that many more steps (unless the caller fixed the step count)
      const int NPA = (p.mode[0].A[2] - 1) / 2;
      pl->M_a = p.mode[0].A[1] * NPA * 2 * p.mode[0].A[0];
      if (h->tune_steps <= 0 && !BRW_ANTI) p.steps_a = p.steps + (p.steps * 53 + 500) / 1000;     // swept 0..17 %: +0.5 % at 5-9 %
    }
    pl->threads = std::min(1024, ((Mmax + 31) / 32) * 32);
    pl->smem = (size_t)p.v_entries * 16 * 8 + (size_t)2 * g.ztot * 4 + 2 * sizeof(BrwStepParams) + 32 * 8 + p.box_sites;
    // offset tables per x-parity of the centre site
    std::vector<int> off(2 * g.ztot);
    for (int par = 0; par < 2; par++)
      for (int k = 0; k < g.ztot; k++) {
        int dx = g.off[k][0], dy = g.off[k][1], dz = g.off[k][2];
        int dxc = g.xs ? ((par + dx) >> 1) : dx;     // arithmetic shift = floor
        int dyc = g.ys ? ((par + dy) >> 1) : dy;     // bcc: y parity == x parity
        off[par * g.ztot + k] = (dz * p.byc + dyc) * p.bxc + dxc;
      }
    // lane-replicated V, layout [shell][centre][nbr][16]
    std::vector<double> vrep((size_t)p.v_entries * 16);
    for (int n = 0; n < g.n_shells; n++) for (int c = 0; c < g.S; c++) for (int s = 0; s < g.S; s++)
      for (int l = 0; l < 16; l++) vrep[(((size_t)(n * g.S + c) * g.S) + s) * 16 + l] = h->hV[(n * g.S + s) * g.S + c];
#define BRW_PLAN_CUDA(x) do { if (brw_cuda_check((x), #x)) { brw_free_plan(pl); return 1; } } while (0)
    BRW_PLAN_CUDA(cudaMalloc(&pl->d_classes, classes.size() * sizeof(int4)));
    BRW_PLAN_CUDA(cudaMalloc(&pl->d_disp, disp.size() * sizeof(int4)));
    BRW_PLAN_CUDA(cudaMalloc(&pl->d_off, off.size() * sizeof(int)));
    BRW_PLAN_CUDA(cudaMalloc(&pl->d_Vrep, vrep.size() * sizeof(double)));
    BRW_PLAN_CUDA(cudaMemcpy(pl->d_classes, classes.data(), classes.size() * sizeof(int4), cudaMemcpyHostToDevice));
    BRW_PLAN_CUDA(cudaMemcpy(pl->d_disp, disp.data(), disp.size() * sizeof(int4), cudaMemcpyHostToDevice));
    BRW_PLAN_CUDA(cudaMemcpy(pl->d_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
    BRW_PLAN_CUDA(cudaMemcpy(pl->d_Vrep, vrep.data(), vrep.size() * sizeof(double), cudaMemcpyHostToDevice));
    if (nbr_swap) BRW_PLAN_CUDA(cudaFuncSetAttribute(brw_box_metropolis_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    else BRW_PLAN_CUDA(cudaFuncSetAttribute(brw_box_metropolis_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    if (!we && !nbr_swap && !h->disable_fast && Mmax <= 768)
      for (const BrwFastEntry &fe : brw_fast_table)
        if (!pl->fast_fn && fe.lat == g.lattice && fe.nsh == g.n_shells && fe.px == p.bxc && fe.py == p.byc &&
            ((Mmax + 31) / 32) * 32 <= fe.maxt) {
          // screening needs <= 5 species (four 8-bit count fields + one inferred) and is a per-handle option
          const bool screen = h->dE_mode >= 1 && g.S <= 5;
          pl->fast_fn = (void *)(screen ? fe.fn_screen : fe.fn);
          pl->screened = screen;
          pl->fast_smem = (size_t)p.v_entries * 16 * 8 + 32 * 8 + (size_t)p.steps * sizeof(BrwStepParams) + p.box_sites;
          pl->threads = std::min(768, ((Mmax + 31) / 32) * 32);
          if (byte_epoch) {
            brw_epoch_tables(h, p, 0.0);
            pl->fast_fn = brw_byte_epoch_kernel_lookup(fe.lat, fe.nsh, fe.px, fe.py, fe.maxt, byte_epoch, h->dE_mode == 0);
            if (!pl->fast_fn) { brw_fail("byte-lattice epoch kernel not instantiated"); brw_free_plan(pl); return 1; }
            pl->screened = h->dE_mode != 0; pl->byte_epoch = true;
            pl->fast_smem = (size_t)p.v_entries * 16 * 8 + 32 * 8 + (size_t)(p.steps / byte_epoch) * sizeof(BrwByteEpochT<4>) +
                            (size_t)((fe.maxt + 31) / 32) * 320 * 4 + (size_t)fe.maxt * 4 +
                            (size_t)(p.box_sites / p.bxc) * brw_byte_epoch_pitch(fe.lat, fe.nsh);
            // programmatic dependent launch only for single-wave grids, and then with a shared-memory request of more than
            // half an SM: CTAs of the next phase launched early must not pile up next to running ones (measured: -20 %)
            int n_sm = 148;
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device);
            if (BRW_BYTE_PDL && (long)p.boxes_per_replica * h->n_replicas <= n_sm) {
              pl->pdl = true;
              pl->fast_smem = std::max(pl->fast_smem, (size_t)116 * 1024);
            }
          }
          BRW_PLAN_CUDA(cudaFuncSetAttribute((const void *)pl->fast_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->fast_smem));
        }
    if (we) {
      // word-lattice kernel: replace the lane-replicated V by its table blob (layout: word_metropolis.cuh)
      const int NL = we->nlimb, NSH = g.n_shells, S = g.S, PAIRS = NSH * NL / 2;
      auto species_of = [](int code) { return code == 4 ? 4 : 3 - code; };
      auto Vn = [&](int n, int centre, int nbr) { return h->hV[(n * S + nbr) * S + centre]; };
      auto U = [&](int sa, int sb, int n, int s) {
        double u = Vn(n, sb, s) - Vn(n, sa, s);
        if (S == 5) u -= Vn(n, sb, 4) - Vn(n, sa, 4);
        return u;
      };
      double umax = 0.0;
      for (int sa = 0; sa < S; sa++) for (int sb = 0; sb < S; sb++) for (int n = 0; n < NSH; n++)
        for (int s2 = 0; s2 < std::min(S, 4); s2++) umax = std::max(umax, std::fabs(U(sa, sb, n, s2)));
      int kexp = 0;
      if (umax > 0.0) kexp = (int)std::floor(std::log2(std::ldexp(1.0, 8 * NL - 2) / umax));
      p.fix_scale = std::ldexp(1.0, -kexp);
      p.row_mul = S <= 4 ? 4 : 5;
      // |fixed-point dE - real dE| <= sum_f |d_f| * 2^-k / 2 <= ztot * 2^-k; the second term covers the f64
      // rounding difference to the reference association as before
      p.guard = g.ztot * p.fix_scale + 1e-9 * g.ztot * umax;
      const int urow_words = (PAIRS + 1) * 64;
      const int tab_words = (urow_words + 2 * g.ztot + 1) & ~1;
      std::vector<int> blob(tab_words + 2 * p.v_entries, 0);
      for (int ca = 0; ca < 5; ca++) for (int cb = 0; cb < 5; cb++) {
        const int sa = species_of(ca), sb = species_of(cb);
        if (sa >= S || sb >= S || sa == sb) continue;
        const int r = ca * p.row_mul + cb;                         // < 32
        auto word = [&](int e) -> int & { return blob[(e / 2) * 64 + r * 2 + (e & 1)]; };
        long long K = 0;
        for (int n = 0; n < NSH; n++)
          for (int s2 = 0; s2 < std::min(S, 4); s2++) {
            long long v = std::llrint(std::ldexp(U(sa, sb, n, s2), kexp));
            K += 128 * v;
            for (int k = 0; k < NL; k++) {
              long long dgt = k == NL - 1 ? v : ((v + 128) & 255) - 128;
              v = (v - dgt) >> 8;
              word(n * NL + k) |= (int)((uint32_t)(uint8_t)(int8_t)dgt << (8 * s2));
            }
          }
        word(2 * PAIRS) = (int)(uint32_t)((unsigned long long)K & 0xFFFFFFFFull);
        word(2 * PAIRS + 1) = (int)(uint32_t)((unsigned long long)K >> 32);
      }
      for (int par = 0; par < 2; par++)
        for (int k = 0; k < g.ztot; k++) {
          int dx = g.off[k][0], dy = g.off[k][1], dz = g.off[k][2];
          int dxc = (par + dx) >> 1, dyc = g.ys ? ((par + dy) >> 1) : dy;
          blob[urow_words + par * g.ztot + k] = dz * we->plp + dyc * we->pxp + dxc;
        }
      if (we->epoch) {
        brw_epoch_tables(h, p, umax);
      }
      std::memcpy(blob.data() + tab_words, h->hV, sizeof(double) * p.v_entries);
      cudaFree(pl->d_Vrep); pl->d_Vrep = nullptr;
      BRW_PLAN_CUDA(cudaMalloc(&pl->d_Vrep, blob.size() * sizeof(int)));
      BRW_PLAN_CUDA(cudaMemcpy(pl->d_Vrep, blob.data(), blob.size() * sizeof(int), cudaMemcpyHostToDevice));
      pl->fast_fn = (void *)(h->dE_mode == 0 ? we->fn_exact : we->fn);
      pl->screened = h->dE_mode != 0; pl->word = true; pl->split = we->split;
      pl->pdl = we->epoch > 0;
      pl->fast_smem = blob.size() * sizeof(int) + 32 * 8 + 2 * (size_t)p.mode[0].A[1] * p.mode[0].A[2] * 4 + 16 +
                      (size_t)(std::max(p.steps, p.steps_a) + 1) * 32 + (size_t)we->plp * p.bzc * 4;
      if (we->epoch)      // epoch table + per-warp count cache instead of the step table
        pl->fast_smem = blob.size() * sizeof(int) + 32 * 8 + 2 * (size_t)p.mode[0].A[1] * p.mode[0].A[2] * 4 + 16 +
                        (size_t)(p.steps / we->epoch) * 32 + 32 * 320 * 4 + (size_t)we->plp * p.bzc * 4;
      pl->threads = 32 * std::min(32, p.mode[0].A[1] * (p.mode[0].A[2] - (we->split ? 1 : 0)));
      p.tma_stages = 0;
      // box load through the bulk-async copy engine (TMA): whole global x-rows of <= 128 bytes land in the storage of the
      // pair-word rows they are expanded into (epoch_metropolis.cuh, brw_pbox_load_tma)
      // (diagnostic switch: BRAWL_CUDA_NO_TMA=1 in the environment keeps the LDG loop -- the tests compare the two loads)
      if (we->epoch && BRW_TMA_LOAD && !getenv("BRAWL_CUDA_NO_TMA") && pl->threads == 32 * we->byc && g.cx % 16 == 0 && g.cx <= 128 &&
          g.n_sites % 16 == 0)
        p.tma_stages = 1;
      BRW_PLAN_CUDA(cudaFuncSetAttribute((const void *)pl->fast_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->fast_smem));
    }
    pl->use_box = true;
    pl->n_slots = p.boxes_per_replica * h->n_replicas;
  } else {
    pl->use_box = false;
    pl->n_slots = h->n_replicas;
  }
  BRW_PLAN_CUDA(cudaMalloc(&pl->d_att, pl->n_slots * sizeof(unsigned long long)));
  BRW_PLAN_CUDA(cudaMalloc(&pl->d_acc, pl->n_slots * sizeof(unsigned long long)));
  BRW_PLAN_CUDA(cudaMalloc(&pl->d_dE, pl->n_slots * sizeof(double)));
  BRW_PLAN_CUDA(cudaMemset(pl->d_att, 0, pl->n_slots * sizeof(unsigned long long)));
  BRW_PLAN_CUDA(cudaMemset(pl->d_acc, 0, pl->n_slots * sizeof(unsigned long long)));
  BRW_PLAN_CUDA(cudaMemset(pl->d_dE, 0, pl->n_slots * sizeof(double)));
  pl->valid = true;
  slot = pl;
  *out = pl;
  return 0;
}

extern "C" int brawl_cuda_metropolis_tune(brawl_cuda_t *h, int bx, int by, int bz, int steps) {
  BRW_ENTER(h);
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  // steps < 0: additionally force the generic (runtime-geometry) kernel -- used by the tests to
  // cross-check the specialised kernels
  h->disable_fast = steps < 0;
  if (steps < 0) steps = -steps - 1;
  h->cubic_period_only = steps >= 100000;      // test hook: steps + 100000 restricts the planner to cubic periods
  if (steps >= 100000) steps -= 100000;
  h->tune_box[0] = bx; h->tune_box[1] = by; h->tune_box[2] = bz; h->tune_steps = steps;
  for (int i = 0; i < 2; i++) { brw_free_plan((BrwPlan *)h->mc_plan[i]); h->mc_plan[i] = nullptr; }
  return 0;
}
extern "C" int brawl_cuda_metropolis_set_mode(brawl_cuda_t *h, int dE_mode) {
  BRW_ENTER(h);
  if (dE_mode < 0 || dE_mode > 2) return brw_fail("dE_mode must be 0 (reference association for every trial), 1 (screened, byte lattice) or 2 (screened, word lattice)");
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  h->dE_mode = dE_mode;
  for (int i = 0; i < 2; i++) { brw_free_plan((BrwPlan *)h->mc_plan[i]); h->mc_plan[i] = nullptr; }
  return 0;
}
extern "C" int brawl_cuda_metropolis_set_layout(brawl_cuda_t *h, int byte_layout_only) {
  BRW_ENTER(h);
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  h->byte_layout = byte_layout_only == 1;
  h->word_split = byte_layout_only == 2 ? 0 : 1;      // 2: word kernels without the two-warp-group split
  // 2, 3: the one-gather-per-step word kernels (3: with the two-warp-group split, the round-1 default); 0: epochs of 4
  // steps (default), 4: epochs of 8, 5: epochs of 2
  h->word_epoch = byte_layout_only == 2 || byte_layout_only == 3 ? 0 : byte_layout_only == 4 ? 8 : byte_layout_only == 5 ? 2 : 4;
  for (int i = 0; i < 2; i++) { brw_free_plan((BrwPlan *)h->mc_plan[i]); h->mc_plan[i] = nullptr; }
  return 0;
}
extern "C" int brawl_cuda_metropolis_plan(brawl_cuda_t *h, int nbr_swap, int *o) {
  BRW_ENTER(h);
  BrwPlan *pl;
  if (brw_build_plan(h, nbr_swap, &pl)) return 1;
  if (o) {
    const BrwBoxMode &m0 = pl->p.mode[0];
    o[0] = pl->use_box; o[1] = m0.P[0] * 10000 + m0.P[1] * 100 + m0.P[2]; o[2] = pl->p.m; o[3] = pl->p.B[0]; o[4] = pl->p.B[1];
    o[5] = pl->p.B[2]; o[6] = pl->Mmax; o[7] = pl->p.boxes_per_replica; o[8] = m0.n_disp; o[9] = pl->p.steps;
    if (pl->fast_fn) o[0] = pl->word ? (pl->screened ? 4 : 5) : pl->byte_epoch ? (pl->screened ? 6 : 7) : pl->screened ? 3 : 2;
    o[0] += 16 * pl->p.n_modes;                 // number of period orientations in bits 4..11
    if (pl->word && pl->split) o[0] += 4096;    // bit 12: two warp groups per CTA, box_z = layer pitch (box depth = pitch + margin)
  }
  return 0;
}

extern "C" int brawl_cuda_metropolis_enqueue(brawl_cuda_t *h, const double *beta, int64_t n_trials, int nbr_swap,
                                             uint64_t seed, uint64_t offset, uint64_t *next_offset,
                                             int64_t *planned, int *n_launches) {
  BRW_ENTER(h);
  if (!beta) return brw_fail("null beta array");
  if (n_trials < 0) return brw_fail("negative trial count");
  BrwPlan *pl;
  if (brw_build_plan(h, nbr_swap, &pl)) return 1;
  if (!h->d_beta) BRW_CUDA(cudaMalloc(&h->d_beta, sizeof(double) * h->n_replicas));
  BRW_CUDA(cudaMemcpyAsync(h->d_beta, beta, sizeof(double) * h->n_replicas, cudaMemcpyHostToDevice, h->stream));
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  int launches = 0;
  int64_t done = 0;
  if (pl->use_box) {
    const BrwBoxParams &p = pl->p;
    // each phase uses one orientation of the period (uniform, from the phase counter); loop until at least
    // n_trials attempts per replica are planned
    int64_t phases = 0;
    while (done < n_trials) {
      uint64_t phase = offset + (uint64_t)phases;
      uint32_t kk1 = k1 ^ (uint32_t)(phase >> 32) * 0x9E3779B9u;
      const int mode = (int)brw_below(brw_philox(0xFFFFFFFDu, 0u, 0u, (uint32_t)phase, k0, kk1).x, (uint32_t)p.n_modes);
      if (pl->fast_fn && pl->pdl) {
        // epoch kernels: programmatic dependent launch -- the next phase's CTAs start their table set-up while the last
        // CTAs of this phase finish (the kernel waits with griddepcontrol.wait before it touches the lattice)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(pl->n_slots); cfg.blockDim = dim3(pl->threads); cfg.dynamicSmemBytes = pl->fast_smem; cfg.stream = h->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        BRW_CUDA(cudaLaunchKernelEx(&cfg, (BrwFastKernel)pl->fast_fn, h->g, p, h->d_lat, (const double *)h->d_beta,
                                    (const double *)pl->d_Vrep, (const int4 *)pl->d_classes, (const int4 *)pl->d_disp, k0, kk1,
                                    (uint32_t)phase, mode, pl->d_att, pl->d_acc, pl->d_dE));
      } else if (pl->fast_fn)
        ((BrwFastKernel)pl->fast_fn)<<<pl->n_slots, pl->threads, pl->fast_smem, h->stream>>>(
            h->g, p, h->d_lat, h->d_beta, pl->d_Vrep, pl->d_classes, pl->d_disp, k0, kk1, (uint32_t)phase, mode, pl->d_att,
            pl->d_acc, pl->d_dE);
      else if (nbr_swap)
        brw_box_metropolis_kernel<1><<<pl->n_slots, pl->threads, pl->smem, h->stream>>>(
            h->g, p, h->d_lat, h->d_beta, pl->d_Vrep, pl->d_off, pl->d_classes, pl->d_disp, k0, kk1, (uint32_t)phase, mode,
            pl->d_att, pl->d_acc, pl->d_dE);
      else
        brw_box_metropolis_kernel<0><<<pl->n_slots, pl->threads, pl->smem, h->stream>>>(
            h->g, p, h->d_lat, h->d_beta, pl->d_Vrep, pl->d_off, pl->d_classes, pl->d_disp, k0, kk1, (uint32_t)phase, mode,
            pl->d_att, pl->d_acc, pl->d_dE);
      BRW_LAUNCH_CHECK("brw_box_metropolis_kernel");
      launches++;
      phases++;
      done += ((int64_t)p.mode[mode].M * p.steps + (int64_t)pl->M_a * (p.steps_a - p.steps)) * p.boxes_per_replica;
    }
    if (next_offset) *next_offset = offset + (uint64_t)phases;
  } else {
    if (n_trials > 0) {
      BrwWalkerLayout lay;
      if (brw_walker_prepare(h->g, 0, brw_chain_metropolis_kernel, lay)) return 1;
      if (h->dE_mode == 0) lay.fast_delta = 0;      // reference association for every trial
      brw_chain_metropolis_kernel<<<(h->n_replicas + BRW_WALKER_WARPS - 1) / BRW_WALKER_WARPS, 32 * BRW_WALKER_WARPS,
                                    lay.total(), h->stream>>>(
          h->g, lay, h->d_V, h->d_lat, h->d_beta, h->n_replicas, (long)n_trials, nbr_swap, k0, k1, (uint32_t)offset,
          (uint32_t)(offset >> 32), pl->d_att, pl->d_acc, pl->d_dE);
      BRW_LAUNCH_CHECK("brw_chain_metropolis_kernel");
      launches++;
    }
    done = n_trials;
    if (next_offset) *next_offset = offset + 1;
  }
  if (planned) *planned = done;
  if (n_launches) *n_launches = launches;
  h->last_plan = nbr_swap ? 1 : 0;
  h->last_launches = launches;
  return 0;
}

extern "C" int brawl_cuda_metropolis_last_launches(brawl_cuda_t *h, int *n) {
  if (!h || !n) return brw_fail("null argument");
  *n = h->last_launches;
  return 0;
}

extern "C" int brawl_cuda_metropolis_counters(brawl_cuda_t *h, int reset, int64_t *att, int64_t *acc, double *dE) {
  BRW_ENTER(h);
  BrwPlan *pl = (BrwPlan *)h->mc_plan[h->last_plan];
  if (!pl || !pl->valid) return brw_fail("no Metropolis run has been enqueued on this handle");
  int n = pl->n_slots, per = n / h->n_replicas;
  std::vector<unsigned long long> a(n), c(n);
  std::vector<double> d(n);
  BRW_CUDA(cudaMemcpyAsync(a.data(), pl->d_att, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaMemcpyAsync(c.data(), pl->d_acc, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaMemcpyAsync(d.data(), pl->d_dE, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (reset) {
    BRW_CUDA(cudaMemsetAsync(pl->d_att, 0, n * sizeof(unsigned long long), h->stream));
    BRW_CUDA(cudaMemsetAsync(pl->d_acc, 0, n * sizeof(unsigned long long), h->stream));
    BRW_CUDA(cudaMemsetAsync(pl->d_dE, 0, n * sizeof(double), h->stream));
  }
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  for (int r = 0; r < h->n_replicas; r++) {
    unsigned long long A = 0, C = 0; double D = 0.0;
    for (int b = 0; b < per; b++) { A += a[r * per + b]; C += c[r * per + b]; D += d[r * per + b]; }
    if (att) att[r] = (int64_t)A;
    if (acc) acc[r] = (int64_t)C;
    if (dE) dE[r] = D;
  }
  return 0;
}

extern "C" int brawl_cuda_metropolis_run(brawl_cuda_t *h, const double *beta, int64_t n_trials, int nbr_swap,
                                         uint64_t seed, uint64_t offset, uint64_t *next_offset, int64_t *att,
                                         int64_t *acc, double *dE) {
  BRW_ENTER(h);
  BrwPlan *pl;
  if (brw_build_plan(h, nbr_swap, &pl)) return 1;
  h->last_plan = nbr_swap ? 1 : 0;
  if (brawl_cuda_metropolis_counters(h, 1, nullptr, nullptr, nullptr)) return 1;   // zero
  if (brawl_cuda_metropolis_enqueue(h, beta, n_trials, nbr_swap, seed, offset, next_offset, nullptr, nullptr)) return 1;
  return brawl_cuda_metropolis_counters(h, 1, att, acc, dE);
}

// ---- SRO ------------------------------------------------------------------------------------------------------
// WC shell radii as lattice_shells computes them (src/analytics.f90:205-275): sorted distinct
// single-precision distances from the origin cell to every lattice site of the box; then every raw
// offset of the reference's cube scan (half-width min(n,5), :336-394) that lands on a site and matches
// a shell radius within 1e-3.
static void brw_sro_offsets(const BrwGeom &g, int wc_range, std::vector<int4> &out, std::vector<double> &radii) {
  std::vector<double> all;
  for (int z = 0; z < g.gz; z++) for (int y = 0; y < g.gy; y++) for (int x = 0; x < g.gx; x++)
    if (brw_is_site(g, x, y, z)) all.push_back((double)sqrtf((float)(z * z) + (float)(y * y) + (float)(x * x)));
  std::sort(all.begin(), all.end());
  radii.clear();
  // the reference's array is padded with zeros, so 0.0 is always the first distinct value
  radii.push_back(0.0);
  for (size_t i = 0; i + 1 < all.size() && (int)radii.size() < wc_range; i++)
    if (std::fabs(all[i] - all[i + 1]) >= 1e-3 && all[i] > radii.back() + 1e-3) radii.push_back(all[i]);
  if ((int)radii.size() < wc_range && !all.empty() && all.back() > radii.back() + 1e-3) { /* last element is never emitted by the reference loop */ }
  while ((int)radii.size() < wc_range) radii.push_back(0.0);   // shells(l) stays 0.0 when the box is too small
  int l1 = std::min(g.gx / 2, 5), l2 = std::min(g.gy / 2, 5), l3 = std::min(g.gz / 2, 5);
  out.clear();
  for (int dz = -l3; dz <= l3; dz++) for (int dy = -l2; dy <= l2; dy++) for (int dx = -l1; dx <= l1; dx++) {
    if (!brw_is_site(g, dx & 1 ? 1 : 0, dy & 1 ? 1 : 0, dz & 1 ? 1 : 0)) continue;   // offset between two sites
    double dist = std::sqrt((double)(dx * dx + dy * dy + dz * dz));
    for (int l = 0; l < wc_range; l++)
      if (std::fabs(dist - radii[l]) < 1e-3) out.push_back(make_int4(dx, dy, dz, l));
  }
}
extern "C" int brawl_cuda_radial_counts(brawl_cuda_t *h, int replica, int wc_range, int64_t *cnt, int64_t *species_count) {
  BRW_ENTER(h);
  BRW_REPLICA(h, replica);
  if (wc_range < 1 || wc_range > 64) return brw_fail("wc_range out of range");
  if (!cnt || !species_count) return brw_fail("null output pointer");
  const BrwGeom &g = h->g;
  std::vector<int4> offs; std::vector<double> radii;
  brw_sro_offsets(g, wc_range, offs, radii);
  size_t n_cnt = (size_t)wc_range * g.S * g.S;
  size_t nb = offs.size() * sizeof(int4) + (n_cnt + g.S + 2) * sizeof(unsigned long long) + 16;
  if (brw_small(h, nb)) return 1;
  unsigned long long *d_cnt = (unsigned long long *)h->d_small, *d_sc = d_cnt + n_cnt;
  int4 *d_off = (int4 *)(d_cnt + ((n_cnt + g.S + 1) & ~(size_t)1));      // 16-byte aligned
  BRW_CUDA(cudaMemsetAsync(d_cnt, 0, (n_cnt + g.S) * sizeof(unsigned long long), h->stream));
  BRW_CUDA(cudaMemcpyAsync(d_off, offs.data(), offs.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
  size_t smem = (n_cnt + g.S) * sizeof(unsigned int);
  brw_radial_counts_kernel<<<grid_for(g.n_sites, 256, 592), 256, smem, h->stream>>>(g, h->d_lat + (size_t)replica * g.n_sites, d_off,
                                                                                     (int)offs.size(), wc_range, d_cnt, d_sc);
  BRW_LAUNCH_CHECK("brw_radial_counts_kernel");
  std::vector<unsigned long long> hc(n_cnt + g.S);
  BRW_CUDA(cudaMemcpyAsync(hc.data(), d_cnt, hc.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i < n_cnt; i++) cnt[i] = (int64_t)hc[i];
  for (int i = 0; i < g.S; i++) species_count[i] = (int64_t)hc[n_cnt + i];
  return 0;
}

// replicas [first, first+n) in one launch (blockIdx.y = replica): cnt[n][wc_range][S][S], species_count[n][S]
extern "C" int brawl_cuda_radial_counts_batch(brawl_cuda_t *h, int first, int n, int wc_range, int64_t *cnt, int64_t *species_count) {
  BRW_ENTER(h);
  if (n < 1 || first < 0 || first + n > h->n_replicas) return brw_fail("replica range [%d,%d) out of [0,%d)", first, first + n, h->n_replicas);
  if (n > 65535) return brw_fail("at most 65535 replicas per radial_counts_batch call");
  if (wc_range < 1 || wc_range > 64) return brw_fail("wc_range out of range");
  if (!cnt || !species_count) return brw_fail("null output pointer");
  const BrwGeom &g = h->g;
  std::vector<int4> offs; std::vector<double> radii;
  brw_sro_offsets(g, wc_range, offs, radii);
  const size_t n_cnt = (size_t)wc_range * g.S * g.S, nh = n_cnt + g.S;
  const size_t cnt_bytes = ((nh * n * sizeof(unsigned long long)) + 15) & ~(size_t)15;
  if (brw_small(h, cnt_bytes + offs.size() * sizeof(int4) + 16)) return 1;
  unsigned long long *d_cnt = (unsigned long long *)h->d_small;
  int4 *d_off = (int4 *)((char *)h->d_small + cnt_bytes);
  BRW_CUDA(cudaMemsetAsync(d_cnt, 0, cnt_bytes, h->stream));
  BRW_CUDA(cudaMemcpyAsync(d_off, offs.data(), offs.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
  const int bx = std::max(1, std::min(grid_for(g.n_sites, 256, 592), (148 * 8 + n - 1) / n));
  brw_radial_counts_kernel<<<dim3(bx, n), 256, nh * sizeof(unsigned int), h->stream>>>(g, h->d_lat + (size_t)first * g.n_sites, d_off,
                                                                                    (int)offs.size(), wc_range, d_cnt, d_cnt + n_cnt);
  BRW_LAUNCH_CHECK("brw_radial_counts_kernel (batch)");
  std::vector<unsigned long long> hc(nh * n);
  BRW_CUDA(cudaMemcpyAsync(hc.data(), d_cnt, hc.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  for (int r = 0; r < n; r++) {
    for (size_t i = 0; i < n_cnt; i++) cnt[(size_t)r * n_cnt + i] = (int64_t)hc[(size_t)r * nh + i];
    for (int i = 0; i < g.S; i++) species_count[(size_t)r * g.S + i] = (int64_t)hc[(size_t)r * nh + n_cnt + i];
  }
  return 0;
}

// ---- Wang-Landau -------------------------------------------------------------------------------------------------
extern "C" int brawl_cuda_wl_sweeps_replay(brawl_cuda_t *h, int replica, double *lng, double *hist, const double *edges,
                                           int bins, int win_lo, int win_hi, double wl_f, int64_t n_trials, int nbr_swap,
                                           uint32_t *mt, int64_t *n_accept, double *e_final) {
  BRW_ENTER(h);
  BRW_REPLICA(h, replica);
  if (!lng || !hist || !edges || !mt) return brw_fail("null array argument");
  if (bins < 1 || win_lo < 1 || win_hi > bins || win_lo > win_hi) return brw_fail("bad energy window [%d,%d] for %d bins", win_lo, win_hi, bins);
  const BrwGeom &g = h->g;
  int nh = win_hi - win_lo + 1;
  size_t nb = sizeof(double) * (bins + nh + 2) + 625 * 4 + 16;
  if (brw_small(h, nb)) return 1;
  double *d_lng = (double *)h->d_small, *d_hist = d_lng + bins, *d_e = d_hist + nh;   // d_e[0]=start, [1]=final
  unsigned long long *d_acc = (unsigned long long *)(d_e + 2);
  uint32_t *d_mt = (uint32_t *)(d_acc + 1);
  // e_unswapped = full_energy(config) at entry (:547), exact order
  if (brw_ensure_scratch(h, (size_t)g.n_sites * sizeof(double))) return 1;
  brw_site_energy_kernel<<<grid_for(g.n_sites, 256), 256, 0, h->stream>>>(g, h->d_V, h->d_lat + (size_t)replica * g.n_sites, h->d_scratch, 1);
  BRW_LAUNCH_CHECK("brw_site_energy_kernel");
  brw_ordered_sum_kernel<<<1, 256, 0, h->stream>>>(h->d_scratch, g.n_sites, d_e);
  BRW_LAUNCH_CHECK("brw_ordered_sum_kernel");
  double e_start = 0.0;
  BRW_CUDA(cudaMemcpyAsync(&e_start, d_e, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaMemcpyAsync(d_lng, lng, sizeof(double) * bins, cudaMemcpyHostToDevice, h->stream));
  BRW_CUDA(cudaMemcpyAsync(d_hist, hist, sizeof(double) * nh, cudaMemcpyHostToDevice, h->stream));
  BRW_CUDA(cudaMemcpyAsync(d_mt, mt, 625 * 4, cudaMemcpyHostToDevice, h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  int hist_every = (int)(0.02 * (double)(float)g.n_sites);          // INT(0.02_real64*REAL(n_atoms)), :605
  double range = edges[bins] - edges[0];
  {
    // the reference indexes wl_logdos(ibin) unguarded; its walkers are inside their window by
    // construction (enter_energy_window, :643-741).  Refuse anything else instead of corrupting memory.
    int ib = (int)(((e_start - edges[0]) / range) * (double)bins) + 1;
    if (!(e_start == e_start) || ib < win_lo || ib > win_hi)   // (energies up to one bin below E_0 map to bin 1, like the reference)
      return brw_fail("walker energy %.10g (bin %d) is outside its window [%d,%d]", e_start, ib, win_lo, win_hi);
  }
  brw_wl_replay_kernel<<<1, 32, 0, h->stream>>>(g, h->d_V, h->d_lat + (size_t)replica * g.n_sites, d_lng, d_hist, edges[0], range, bins,
                                                win_lo, win_hi, wl_f, (long)n_trials, hist_every, nbr_swap, e_start, d_mt, d_acc, d_e + 1);
  BRW_LAUNCH_CHECK("brw_wl_replay_kernel");
  unsigned long long acc = 0; double ef = 0.0;
  BRW_CUDA(cudaMemcpyAsync(lng, d_lng, sizeof(double) * bins, cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaMemcpyAsync(hist, d_hist, sizeof(double) * nh, cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaMemcpyAsync(mt, d_mt, 625 * 4, cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaMemcpyAsync(&acc, d_acc, sizeof acc, cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaMemcpyAsync(&ef, d_e + 1, sizeof ef, cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  if (n_accept) *n_accept = (int64_t)acc;
  if (e_final) *e_final = ef;
  return 0;
}

extern "C" int brawl_cuda_wl_enter_window_replay(brawl_cuda_t *h, int replica, double e_start, double target, double lo_e,
                                                 double hi_e, double two_sigma_sq, int64_t period, int64_t *i_steps_io,
                                                 int64_t max_iters, int resume, uint32_t *mt, double *e_out, int *status,
                                                 int64_t *iters_begun) {
  BRW_ENTER(h);
  BRW_REPLICA(h, replica);
  if (!mt || !i_steps_io || !e_out || !status || !iters_begun) return brw_fail("null argument");
  if (period < 1 || max_iters < 0 || !(two_sigma_sq > 0.0)) return brw_fail("bad window-entry parameters");
  const BrwGeom &g = h->g;
  if (brw_small(h, 625 * 4 + 64)) return 1;
  double *d_e = (double *)h->d_small;
  long *d_out = (long *)(d_e + 1);
  uint32_t *d_mt = (uint32_t *)(d_out + 3);
  BRW_CUDA(cudaMemcpyAsync(d_mt, mt, 625 * 4, cudaMemcpyHostToDevice, h->stream));
  brw_wl_enter_replay_kernel<<<1, 32, 0, h->stream>>>(g, h->d_V, h->d_lat + (size_t)replica * g.n_sites, e_start, target, lo_e, hi_e,
                                                      two_sigma_sq, (long)period, (long)*i_steps_io, (long)max_iters, resume, d_mt,
                                                      d_e, d_out);
  BRW_LAUNCH_CHECK("brw_wl_enter_replay_kernel");
  long o3[3] = {0, 0, 0};
  BRW_CUDA(cudaMemcpyAsync(mt, d_mt, 625 * 4, cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaMemcpyAsync(e_out, d_e, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaMemcpyAsync(o3, d_out, sizeof(o3), cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  *status = (int)o3[0]; *iters_begun = (int64_t)o3[1]; *i_steps_io = (int64_t)o3[2];
  return 0;
}

extern "C" int brawl_cuda_wl_sweeps(brawl_cuda_t *h, int n_walkers, double *lng, double *hist, int on_device,
                                    const double *edges, int bins, const int32_t *win_lo, const int32_t *win_hi,
                                    int hist_stride, double wl_f, int64_t n_trials, int nbr_swap, uint64_t seed,
                                    uint64_t offset, int64_t *n_accept, double *e_final) {
  BRW_ENTER(h);
  if (n_walkers < 1 || n_walkers > h->n_replicas) return brw_fail("n_walkers %d out of 1..%d", n_walkers, h->n_replicas);
  if (!lng || !hist || !edges || !win_lo || !win_hi) return brw_fail("null array argument");
  for (int w = 0; w < n_walkers; w++)
    if (win_lo[w] < 1 || win_hi[w] > bins || win_lo[w] > win_hi[w] || win_hi[w] - win_lo[w] + 1 > hist_stride)
      return brw_fail("bad energy window [%d,%d] for walker %d", win_lo[w], win_hi[w], w);
  const BrwGeom &g = h->g;
  size_t n_l = (size_t)n_walkers * bins, n_h = (size_t)n_walkers * hist_stride;
  size_t nb = sizeof(double) * (on_device ? 0 : n_l + n_h) + sizeof(double) * n_walkers + sizeof(unsigned long long) * n_walkers + 2 * sizeof(int) * n_walkers + 64;
  if (brw_small(h, nb)) return 1;
  double *d_e = (double *)h->d_small;
  unsigned long long *d_acc = (unsigned long long *)(d_e + n_walkers);
  int *d_lo = (int *)(d_acc + n_walkers), *d_hi = d_lo + n_walkers;
  double *d_lng = on_device ? lng : (double *)(d_hi + n_walkers + (n_walkers & 1)), *d_hist = on_device ? hist : d_lng + n_l;
  if (!on_device) {
    BRW_CUDA(cudaMemcpyAsync(d_lng, lng, sizeof(double) * n_l, cudaMemcpyHostToDevice, h->stream));
    BRW_CUDA(cudaMemcpyAsync(d_hist, hist, sizeof(double) * n_h, cudaMemcpyHostToDevice, h->stream));
  }
  BRW_CUDA(cudaMemcpyAsync(d_lo, win_lo, sizeof(int) * n_walkers, cudaMemcpyHostToDevice, h->stream));
  BRW_CUDA(cudaMemcpyAsync(d_hi, win_hi, sizeof(int) * n_walkers, cudaMemcpyHostToDevice, h->stream));
  if (brw_total_energy_dev(h, 0, n_walkers, 1, d_e)) return 1;                    // :547
  int hist_every = (int)(0.02 * (double)(float)g.n_sites);
  double range = edges[bins] - edges[0];
  {
    std::vector<double> e0(n_walkers);
    BRW_CUDA(cudaMemcpyAsync(e0.data(), d_e, sizeof(double) * n_walkers, cudaMemcpyDeviceToHost, h->stream));
    BRW_CUDA(cudaStreamSynchronize(h->stream));
    for (int w = 0; w < n_walkers; w++) {
      int ib = (int)(((e0[w] - edges[0]) / range) * (double)bins) + 1;
      if (!(e0[w] == e0[w]) || ib < win_lo[w] || ib > win_hi[w])
        return brw_fail("walker %d energy %.10g (bin %d) is outside its window [%d,%d]", w, e0[w], ib, win_lo[w], win_hi[w]);
    }
  }
  BrwWalkerLayout lay;
  if (brw_walker_prepare(g, bins + hist_stride, brw_wl_walker_kernel, lay)) return 1;
  if (h->dE_mode == 0) lay.fast_delta = 0;      // reference association for every trial
  brw_wl_walker_kernel<<<(n_walkers + BRW_WALKER_WARPS - 1) / BRW_WALKER_WARPS, 32 * BRW_WALKER_WARPS, lay.total(), h->stream>>>(g, lay, h->d_V, h->d_lat, d_lng, d_hist, edges[0], range, bins, d_lo, d_hi,
                                                                    hist_stride, wl_f, (long)n_trials, hist_every, nbr_swap, (uint32_t)seed,
                                                                    (uint32_t)(seed >> 32), (uint32_t)offset, (uint32_t)(offset >> 32),
                                                                    n_walkers, d_e, d_acc);
  BRW_LAUNCH_CHECK("brw_wl_walker_kernel");
  std::vector<unsigned long long> acc(n_walkers);
  if (!on_device) {
    BRW_CUDA(cudaMemcpyAsync(lng, d_lng, sizeof(double) * n_l, cudaMemcpyDeviceToHost, h->stream));
    BRW_CUDA(cudaMemcpyAsync(hist, d_hist, sizeof(double) * n_h, cudaMemcpyDeviceToHost, h->stream));
  }
  if (n_accept) BRW_CUDA(cudaMemcpyAsync(acc.data(), d_acc, sizeof(unsigned long long) * n_walkers, cudaMemcpyDeviceToHost, h->stream));
  if (e_final) BRW_CUDA(cudaMemcpyAsync(e_final, d_e, sizeof(double) * n_walkers, cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  if (n_accept) for (int w = 0; w < n_walkers; w++) n_accept[w] = (int64_t)acc[w];
  return 0;
}

extern "C" int brawl_cuda_wl_enter_window(brawl_cuda_t *h, int n_walkers, const double *target, const double *lo_e,
                                          const double *hi_e, double inv_two_sigma_sq, int64_t max_trials, uint64_t seed,
                                          uint64_t offset, double *energies, int32_t *entered) {
  BRW_ENTER(h);
  if (n_walkers < 1 || n_walkers > h->n_replicas) return brw_fail("n_walkers %d out of 1..%d", n_walkers, h->n_replicas);
  if (!target || !lo_e || !hi_e || !energies || !entered) return brw_fail("null array argument");
  size_t nb = (size_t)n_walkers * (4 * sizeof(double) + sizeof(int)) + 64;
  if (brw_small(h, nb)) return 1;
  double *d_e = (double *)h->d_small, *d_t = d_e + n_walkers, *d_lo = d_t + n_walkers, *d_hi = d_lo + n_walkers;
  int *d_ent = (int *)(d_hi + n_walkers);
  BRW_CUDA(cudaMemcpyAsync(d_t, target, sizeof(double) * n_walkers, cudaMemcpyHostToDevice, h->stream));
  BRW_CUDA(cudaMemcpyAsync(d_lo, lo_e, sizeof(double) * n_walkers, cudaMemcpyHostToDevice, h->stream));
  BRW_CUDA(cudaMemcpyAsync(d_hi, hi_e, sizeof(double) * n_walkers, cudaMemcpyHostToDevice, h->stream));
  if (brw_total_energy_dev(h, 0, n_walkers, 1, d_e)) return 1;                    // :658
  BrwWalkerLayout lay;
  if (brw_walker_prepare(h->g, 0, brw_wl_enter_window_kernel, lay)) return 1;
  if (h->dE_mode == 0) lay.fast_delta = 0;      // reference association for every trial
  brw_wl_enter_window_kernel<<<(n_walkers + BRW_WALKER_WARPS - 1) / BRW_WALKER_WARPS, 32 * BRW_WALKER_WARPS, lay.total(), h->stream>>>(
      h->g, lay, h->d_V, h->d_lat, d_t, d_lo, d_hi, inv_two_sigma_sq, (long)max_trials, (uint32_t)seed, (uint32_t)(seed >> 32),
      (uint32_t)offset, (uint32_t)(offset >> 32), n_walkers, d_e, d_ent);
  BRW_LAUNCH_CHECK("brw_wl_enter_window_kernel");
  BRW_CUDA(cudaMemcpyAsync(energies, d_e, sizeof(double) * n_walkers, cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaMemcpyAsync(entered, d_ent, sizeof(int) * n_walkers, cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int brawl_cuda_lattice_ptr(brawl_cuda_t *h, void **dev_ptr, int64_t *bytes_per_replica) {
  if (!h) return brw_fail("null handle");
  if (dev_ptr) *dev_ptr = h->d_lat;
  if (bytes_per_replica) *bytes_per_replica = h->g.n_sites;
  return 0;
}

__global__ void brw_swap_replicas_kernel(uint8_t *a, uint8_t *b, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint8_t t = a[i]; a[i] = b[i]; b[i] = t;
  }
}
extern "C" int brawl_cuda_swap_replicas(brawl_cuda_t *h, int a, int b) {
  BRW_ENTER(h);
  BRW_REPLICA(h, a); BRW_REPLICA(h, b);
  if (a == b) return 0;
  brw_swap_replicas_kernel<<<grid_for(h->g.n_sites, 256, 148), 256, 0, h->stream>>>(
      h->d_lat + (size_t)a * h->g.n_sites, h->d_lat + (size_t)b * h->g.n_sites, h->g.n_sites);
  BRW_LAUNCH_CHECK("brw_swap_replicas_kernel");
  return 0;
}

extern "C" int brawl_cuda_wl_window_average(brawl_cuda_t *h, double *dev_array, int len, int walkers_per_window, int n_windows, double divisor) {
  BRW_ENTER(h);
  if (!dev_array || len < 1 || walkers_per_window < 1 || n_windows < 1) return brw_fail("bad window-average arguments");
  brw_wl_window_average_kernel<<<(n_windows * len + 127) / 128, 128, 0, h->stream>>>(dev_array, len, walkers_per_window, n_windows, divisor);
  BRW_LAUNCH_CHECK("brw_wl_window_average_kernel");
  return 0;
}

// ---- nested sampling ------------------------------------------------------------------------------------------------
extern "C" int brawl_cuda_ns_walk_replay(brawl_cuda_t *h, int replica, double *energy, double e_limit, int64_t n_steps,
                                         uint32_t *mt, int64_t *n_accept) {
  BRW_ENTER(h);
  BRW_REPLICA(h, replica);
  if (!energy || !mt) return brw_fail("null argument");
  if (brw_small(h, 625 * 4 + 32)) return 1;
  double *d_e = (double *)h->d_small;
  unsigned long long *d_acc = (unsigned long long *)(d_e + 1);
  uint32_t *d_mt = (uint32_t *)(d_acc + 1);
  BRW_CUDA(cudaMemcpyAsync(d_e, energy, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  BRW_CUDA(cudaMemcpyAsync(d_mt, mt, 625 * 4, cudaMemcpyHostToDevice, h->stream));
  brw_ns_replay_kernel<<<1, 32, 0, h->stream>>>(h->g, h->d_V, h->d_lat + (size_t)replica * h->g.n_sites, d_e, e_limit, (long)n_steps, d_mt, d_acc);
  BRW_LAUNCH_CHECK("brw_ns_replay_kernel");
  unsigned long long acc = 0;
  BRW_CUDA(cudaMemcpyAsync(energy, d_e, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaMemcpyAsync(mt, d_mt, 625 * 4, cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaMemcpyAsync(&acc, d_acc, sizeof acc, cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  if (n_accept) *n_accept = (int64_t)acc;
  return 0;
}
extern "C" int brawl_cuda_ns_walk(brawl_cuda_t *h, int n_walkers, const int32_t *ids, double *energies, const double *e_limit,
                                  int64_t n_steps, uint64_t seed, uint64_t offset, int64_t *n_accept) {
  BRW_ENTER(h);
  if (n_walkers < 1) return brw_fail("n_walkers must be >= 1");
  if (!ids || !energies || !e_limit) return brw_fail("null array argument");
  std::vector<char> used(h->n_replicas, 0);
  for (int w = 0; w < n_walkers; w++) {
    if (ids[w] < 0 || ids[w] >= h->n_replicas) return brw_fail("walker id %d out of range", ids[w]);
    if (used[ids[w]]) return brw_fail("walker id %d listed twice", ids[w]);
    used[ids[w]] = 1;
  }
  size_t nb = (size_t)n_walkers * (2 * sizeof(double) + sizeof(unsigned long long) + sizeof(int)) + 64;
  if (brw_small(h, nb)) return 1;
  double *d_e = (double *)h->d_small, *d_lim = d_e + n_walkers;
  unsigned long long *d_acc = (unsigned long long *)(d_lim + n_walkers);
  int *d_ids = (int *)(d_acc + n_walkers);
  BRW_CUDA(cudaMemcpyAsync(d_e, energies, sizeof(double) * n_walkers, cudaMemcpyHostToDevice, h->stream));
  BRW_CUDA(cudaMemcpyAsync(d_lim, e_limit, sizeof(double) * n_walkers, cudaMemcpyHostToDevice, h->stream));
  BRW_CUDA(cudaMemcpyAsync(d_ids, ids, sizeof(int) * n_walkers, cudaMemcpyHostToDevice, h->stream));
  BrwWalkerLayout lay;
  if (brw_walker_prepare(h->g, 0, brw_ns_walker_kernel, lay)) return 1;
  if (h->dE_mode == 0) lay.fast_delta = 0;      // reference association for every trial
  brw_ns_walker_kernel<<<(n_walkers + BRW_WALKER_WARPS - 1) / BRW_WALKER_WARPS, 32 * BRW_WALKER_WARPS, lay.total(), h->stream>>>(h->g, lay, h->d_V, h->d_lat, d_ids, d_e, d_lim, (long)n_steps, (uint32_t)seed,
                                                                    (uint32_t)(seed >> 32), (uint32_t)offset, (uint32_t)(offset >> 32),
                                                                    n_walkers, d_acc);
  BRW_LAUNCH_CHECK("brw_ns_walker_kernel");
  std::vector<unsigned long long> acc(n_walkers);
  BRW_CUDA(cudaMemcpyAsync(energies, d_e, sizeof(double) * n_walkers, cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaMemcpyAsync(acc.data(), d_acc, sizeof(unsigned long long) * n_walkers, cudaMemcpyDeviceToHost, h->stream));
  BRW_CUDA(cudaStreamSynchronize(h->stream));
  if (n_accept) for (int w = 0; w < n_walkers; w++) n_accept[w] = (int64_t)acc[w];
  return 0;
}

#include "wl_resident.inc"
