// Microbenchmark: latency of the pair-word neighbour-count gather (pair_gather.inc) for one warp, as in
// brw_box_metropolis_word_kernel: 28 active lanes (14 A sites at even words, 14 B sites at odd words of another row),
// box pitch 32 / 1024 words, both sites of a trial, nibble->byte expansion and the D words.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I brawl_b200/csrc -o tools/micro/gather_lat tools/micro/gather_lat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t brw_nib2byte(uint32_t x) {
  const uint32_t t = __byte_perm(x, 0u, 0x4140);
  return (t | (t << 4)) & 0x0F0F0F0Fu;
}
__device__ __forceinline__ uint32_t brw_lo16(const uint32_t *p) { return *reinterpret_cast<const uint16_t *>(p); }
#include "pair_gather.inc"
__global__ void k(uint32_t *out, long long *cyc, int iters, int mode) {
  extern __shared__ uint32_t box[];                      // 32 planes x 1024 words
  for (int i = threadIdx.x; i < 32 * 1024; i += blockDim.x) box[i] = 0x00010001u << (4 * ((i * 7) & 3));
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane >= 14, li = lane - 14 * sub;
  const bool active = lane < 28;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (active) {
      // a data-dependent (but in-range) centre so that iterations form one dependent chain like the real steps
      const int row = 4 + ((acc + warp) & 7), pl = 6 + ((acc >> 3) & 15);
      const uint32_t *w1 = box + pl * 1024 + (row + sub) * 32 + 2 + 2 * li + sub;
      const uint32_t *w2 = box + (pl + 2) * 1024 + (row + 12 + sub) * 32 + 2 + 2 * li + sub;
      uint32_t C1[4], C2[4];
      if (mode & 1) BrwPairGatherBcc<4, 32, 1024, 1>::run(w1, C1); else BrwPairGatherBcc<4, 32, 1024, 0>::run(w1, C1);
      if (mode & 2) BrwPairGatherBcc<4, 32, 1024, 1>::run(w2, C2); else BrwPairGatherBcc<4, 32, 1024, 0>::run(w2, C2);
      uint32_t d = 0;
#pragma unroll
      for (int n = 0; n < 4; n++) d ^= 0x80808080u + C1[n] - C2[n];
      acc += d & 1;
    }
    __syncwarp();
  }
  long long t1 = clock64();
  if (lane == 0) cyc[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  uint32_t *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 32 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024 * 4);
  const int iters = 2000;
  for (int warps : {1, 2, 4, 8, 16, 30}) {
    for (int rep = 0; rep < 2; rep++) { k<<<148, 32 * warps, 32 * 1024 * 4>>>(out, cyc, iters, 1); cudaDeviceSynchronize(); }
    long long h[32];
    cudaMemcpy(h, cyc, sizeof(long long) * warps, cudaMemcpyDeviceToHost);
    printf("two-site gather (60 loads + sums + expansions)  warps/SM %2d : %7.1f cycles per trial per warp  (%s)\n", warps,
           (double)h[0] / iters, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
