// Micro-benchmark (not product code): tcgen05.ld throughput per SM as a function of the number of reading warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I transferable3d_b200/csrc tools/ubench/ldtm_bw.cu -o scratch_ab/ldtm_bw
#include <cstdio>
#include "common.cuh"
using namespace t3d;

template <int DEPTH>
__global__ void __launch_bounds__(1024, 1) ldtm_kernel(int iters, unsigned long long* out, unsigned* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t v[DEPTH][32];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) tmem_ld32(base + (uint32_t)(((i * DEPTH + d + warp) * 32) & 511), v[d]);
    tmem_ld_wait();
#pragma unroll
    for (int d = 0; d < DEPTH; ++d)
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[d][j];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

int main() {
  unsigned long long* out; unsigned* sink;
  cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int depth = 1; depth <= 4; depth *= 2)
    for (int warps = 4; warps <= 32; warps *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (depth == 1) ldtm_kernel<1><<<148, warps * 32>>>(iters, out, sink);
        else if (depth == 2) ldtm_kernel<2><<<148, warps * 32>>>(iters, out, sink);
        else ldtm_kernel<4><<<148, warps * 32>>>(iters, out, sink);
        cudaDeviceSynchronize();
      }
      unsigned long long h[148]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      const double bytes = (double)warps * iters * depth * 4096.0;
      printf("depth %d warps %2d: %8llu cycles, %.1f B/clk/SM  (%s)\n", depth, warps, h[0], bytes / (double)h[0], cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
