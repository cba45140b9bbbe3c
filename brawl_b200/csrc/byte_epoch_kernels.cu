// byte_epoch_kernels.cu -- second translation unit of libbrawl_cuda.so: the instantiations of the byte-lattice epoch
// kernels (epoch_byte_metropolis.cuh: fcc 4 / 6 shells, bcc 4 / 6 shells; epochs of 4 / 8 steps; screened / reference
// association; CTAs of <= 512 / 768 threads) behind one lookup function, so that they compile in parallel with
// brawl_cuda.cu.  No device linking: kernels are launched through the returned function pointers.
#include <array>
#include <cstring>
#include <cmath>
#include <vector>
#include "brawl_common.cuh"
#define BRW_TABLE_QUAL static constexpr
#include "shell_tables.inc"
#include "epoch_byte_metropolis.cuh"

typedef void (*BrwFastKernel)(BrwGeom, BrwBoxParams, uint8_t *, const double *, const double *, const int4 *,
                              const int4 *, uint32_t, uint32_t, uint32_t, int, unsigned long long *, unsigned long long *,
                              double *);
namespace {
struct Entry { int lat, nsh, px, py, maxt; BrwFastKernel fn[2][2]; };
#define BRW_BYTE_EPOCH(LAT, NSH, PX, PY, MAXT) {LAT, NSH, PX, PY, MAXT, \
    {{brw_box_metropolis_byte_epoch_kernel<LAT, NSH, PX, PY, false, 4, MAXT>, brw_box_metropolis_byte_epoch_kernel<LAT, NSH, PX, PY, true, 4, MAXT>}, \
     {brw_box_metropolis_byte_epoch_kernel<LAT, NSH, PX, PY, false, 8, MAXT>, brw_box_metropolis_byte_epoch_kernel<LAT, NSH, PX, PY, true, 8, MAXT>}}}
const Entry table[] = {
    BRW_BYTE_EPOCH(1, 4, 32, 32, 512), BRW_BYTE_EPOCH(1, 4, 32, 32, 768), BRW_BYTE_EPOCH(1, 6, 32, 32, 512), BRW_BYTE_EPOCH(1, 6, 32, 32, 768),
    BRW_BYTE_EPOCH(2, 4, 32, 64, 512), BRW_BYTE_EPOCH(2, 4, 32, 64, 768), BRW_BYTE_EPOCH(2, 6, 32, 64, 512), BRW_BYTE_EPOCH(2, 6, 32, 64, 768),
};
}  // namespace

// bytes between consecutive x-rows of the shared-memory box of these kernels (the host sizes the launch with it)
int brw_byte_epoch_pitch(int lat, int nsh) {
  return lat == 1 ? (nsh <= 4 ? BrwBytePitch<1, 4>::value : BrwBytePitch<1, 6>::value)
                  : (nsh <= 4 ? BrwBytePitch<2, 4>::value : BrwBytePitch<2, 6>::value);
}

// epoch_k = 4 or 8; exact != 0: reference association for every trial.  nullptr if not instantiated.
void *brw_byte_epoch_kernel_lookup(int lat, int nsh, int px, int py, int maxt, int epoch_k, int exact) {
  for (const Entry &e : table)
    if (e.lat == lat && e.nsh == nsh && e.px == px && e.py == py && e.maxt == maxt) return (void *)e.fn[epoch_k == 8][exact != 0];
  return nullptr;
}
