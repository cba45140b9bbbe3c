// Microbenchmark: cost of a burst of independent LDS.32 (immediate offsets, one base register) issued by one warp, as a
// function of the number of warps per SM doing the same.  Mirrors the gather of brw_box_metropolis_word_kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/lds_issue tools/micro/lds_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int N> __device__ __forceinline__ unsigned burst(const unsigned *p) {
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < N; i++) s += p[(i * 37) & 1023];       // immediates, conflict-free (lanes 2 words apart below)
  return s;
}
template <int N>
__global__ void k(unsigned *out, long long *cyc, int iters) {
  __shared__ unsigned sm[8192];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 2654435761u;
  __syncthreads();
  const unsigned *p = sm + 2 * (threadIdx.x & 31) + 64 * (threadIdx.x >> 5);
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    acc += burst<N>(p + (acc & 1));
  }
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int N> void run(int warps) {
  unsigned *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 32 * 8);
  const int iters = 2000;
  k<N><<<148, 32 * warps>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  k<N><<<148, 32 * warps>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[32];
  cudaMemcpy(h, cyc, sizeof(long long) * warps, cudaMemcpyDeviceToHost);
  double c = (double)h[0] / iters;
  printf("burst %2d LDS  warps/SM %2d : %7.1f cycles per burst = %5.2f per LDS per warp, SM rate %.2f LDS/cycle\n", N, warps, c, c / N,
         warps * N / c);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 2, 4, 8, 16, 30}) run<30>(w);
  for (int w : {1, 8, 30}) run<8>(w);
  for (int w : {1, 8, 30}) run<60>(w);
  return 0;
}
