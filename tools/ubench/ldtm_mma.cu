// Micro-benchmark (not product code): does tcgen05.ld traffic slow tcgen05.mma down (shared TMEM bandwidth)?
// One CTA per SM: warp 0 issues NMMA back-to-back M=128 x N x K=16 bf16 MMAs (operands = whatever is in shared memory) into
// columns [0, N); warps 4.. read columns [256, 512) with tcgen05.ld.32x32b.x32.  Reports cycles of each alone and together.
#include <cstdio>
#include "common.cuh"
using namespace t3d;

__global__ void __launch_bounds__(1024, 1) k(int nmma, int ncols, int ld_iters, int do_mma, int do_ld, unsigned long long* out, unsigned* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar = sbase + 49152, slot = sbase + 49152 + 16;
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<512>(slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(smem + 49152 + 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (do_mma) {
      const uint32_t idesc = make_idesc_bf16(128, ncols);
      const uint64_t ad = make_sdesc_k128(sbase), bd = make_sdesc_k128(sbase + 16384);
      t0 = clock64();
      for (int i = 0; i < nmma; ++i) umma_bf16_w(tm, ad + 2u * (i & 3), bd + 2u * (i & 3), idesc, i != 0);
      umma_commit_w(bar);
      mbar_wait_w(bar, 0);
      t1 = clock64();
      if (lane == 0) out[blockIdx.x * 2] = (unsigned long long)(t1 - t0);
    }
  } else if (warp >= 4 && do_ld) {
    const uint32_t base = tm + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    t0 = clock64();
    for (int i = 0; i < ld_iters; ++i) {
      uint32_t v[2][32];
      tmem_ld32(base + (uint32_t)(((2 * i + warp) * 32) & 255), v[0]);
      tmem_ld32(base + (uint32_t)(((2 * i + 1 + warp) * 32) & 255), v[1]);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[0][j] ^ v[1][j];
    }
    t1 = clock64();
    if (warp == 4 && lane == 0) out[blockIdx.x * 2 + 1] = (unsigned long long)(t1 - t0);
  }
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tm);
}

int main() {
  unsigned long long* out; unsigned* sink;
  cudaMalloc(&out, 148 * 16); cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  const int nmma = 4000;
  for (int ncols = 128; ncols <= 256; ncols *= 2)
    for (int ldw = 4; ldw <= 16; ldw *= 2) {
      const int ld_iters = 2000 * 8 / ldw;
      for (int mode = 1; mode <= 3; ++mode) {
        cudaMemset(out, 0, 148 * 16);
        for (int rep = 0; rep < 2; ++rep) { k<<<148, (4 + ldw) * 32, 60000>>>(nmma, ncols, ld_iters, mode & 1, mode >> 1, out, sink); cudaDeviceSynchronize(); }
        unsigned long long h[2]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("N=%d ld warps %2d mode %s: mma %.1f cyc/instr (floor %d)   ldtm %.1f B/clk/SM   (%s)\n", ncols, ldw,
               mode == 1 ? "mma only" : mode == 2 ? "ld only " : "both    ", h[0] / (double)nmma, ncols / 2,
               h[1] ? (double)ldw * ld_iters * 8192.0 / (double)h[1] : 0.0, cudaGetErrorString(cudaGetLastError()));
      }
    }
  return 0;
}
