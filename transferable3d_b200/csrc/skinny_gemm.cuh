// HBM-bound GEMM shapes of the training steps that no tile kernel fits: the first layer of every point network has
// K = 3 or 12 input channels over B*N rows.  Its forward (K <= 16), the input gradient of the frozen BoxPC branch
// (N <= 16) and its weight gradient (M <= 16, K = B*N) move ~270 MB for ~0.2 - 6 GFLOP; the generic 64 x 64 CUDA-core
// kernel took 140 - 340 us for each (ncu launch list of the cfg5 step) against a ~45 us HBM floor.  One purpose-built
// kernel per shape, each streaming the big operand exactly once with 128-bit accesses, fp32 FMAs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "simt_ops.cuh"     // apply_act

namespace t3d {

constexpr int kSkinnyMax = 16;
constexpr size_t kSkinnySmemMax = 48 * 1024;   // dynamic shared memory without an opt-in; larger problems take the tiled GEMMs

// (a) C[M,N] = A[M,K] . B[K,N] (+ bias), K <= 16, N % 4 == 0, A row-major (lda), B row-major (ldb), C row-major (ldc).
//     thread = 4 consecutive columns of one row; a block covers 256 * 4 / N rows per step.
__global__ void __launch_bounds__(256) skinny_k_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B,
                                                      long long ldb, const float* __restrict__ bias, float* __restrict__ C,
                                                      long long ldc, int M, int N, int K, int act = 0, float* __restrict__ st_sum = nullptr,
                                                      float* __restrict__ st_sq = nullptr, const float* __restrict__ st_shift = nullptr) {
  extern __shared__ float sB[];                      // [K][N] + bias [N] (+ column sums / sums of squares [2][N] when st_sum)
  for (int i = threadIdx.x; i < K * N; i += 256) sB[i] = B[(long long)(i / N) * ldb + (i % N)];
  for (int i = threadIdx.x; i < N; i += 256) sB[K * N + i] = bias ? bias[i] : 0.0f;
  const bool stats = st_sum != nullptr;
  if (stats) for (int i = threadIdx.x; i < 2 * N; i += 256) sB[K * N + N + i] = 0.0f;
  __syncthreads();
  const int n4 = N / 4, rows_per_step = 256 / n4;
  const int cq = (threadIdx.x % n4) * 4, rl = threadIdx.x / n4;
  if (rl < rows_per_step) {
    // training-mode BN statistics of the output (fused): this thread's 4 columns over all its rows, in registers
    float4 sf = make_float4(0.f, 0.f, 0.f, 0.f), s1 = sf, s2 = sf;
    if (stats && st_shift) sf = make_float4(st_shift[cq], st_shift[cq + 1], st_shift[cq + 2], st_shift[cq + 3]);
    for (long long m = (long long)blockIdx.x * rows_per_step + rl; m < M; m += (long long)gridDim.x * rows_per_step) {
      const float* a = A + m * lda;
      float4 acc = *reinterpret_cast<const float4*>(&sB[K * N + cq]);
      for (int k = 0; k < K; ++k) {
        const float x = __ldg(a + k);
        const float4 w = *reinterpret_cast<const float4*>(&sB[k * N + cq]);
        acc.x = fmaf(x, w.x, acc.x); acc.y = fmaf(x, w.y, acc.y); acc.z = fmaf(x, w.z, acc.z); acc.w = fmaf(x, w.w, acc.w);
      }
      if (act != 0) { acc.x = apply_act(acc.x, act); acc.y = apply_act(acc.y, act); acc.z = apply_act(acc.z, act); acc.w = apply_act(acc.w, act); }
      *reinterpret_cast<float4*>(C + m * ldc + cq) = acc;
      if (stats) {
        const float d0 = acc.x - sf.x, d1 = acc.y - sf.y, d2 = acc.z - sf.z, d3 = acc.w - sf.w;
        s1.x += d0; s1.y += d1; s1.z += d2; s1.w += d3;
        s2.x = fmaf(d0, d0, s2.x); s2.y = fmaf(d1, d1, s2.y); s2.z = fmaf(d2, d2, s2.z); s2.w = fmaf(d3, d3, s2.w);
      }
    }
    if (stats) {
      float* q = &sB[K * N + N + cq];
      atomicAdd(q, s1.x); atomicAdd(q + 1, s1.y); atomicAdd(q + 2, s1.z); atomicAdd(q + 3, s1.w);
      q += N;
      atomicAdd(q, s2.x); atomicAdd(q + 1, s2.y); atomicAdd(q + 2, s2.z); atomicAdd(q + 3, s2.w);
    }
  }
  if (stats) {
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += 256) { atomicAdd(st_sum + i, sB[K * N + N + i]); atomicAdd(st_sq + i, sB[K * N + 2 * N + i]); }
  }
}

// y0[n] = sum_k a(k) B[k, n] + bias[n] for one row a, a(k) = relu(a_scale[k] * a[k] + a_shift[k]) when a_scale != null.
// Block = 32 columns x 8 k-slices (a serial loop over K per column took 30 us at K = 512: a dependent chain of L2 round trips).
__global__ void __launch_bounds__(256) row0_kernel(const float* __restrict__ a, const float* __restrict__ a_scale,
                                                  const float* __restrict__ a_shift, const float* __restrict__ B, int ldb,
                                                  const float* __restrict__ bias, int K, int N, float* __restrict__ y0) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31), ks = threadIdx.x >> 5;
  float acc = 0.0f;
  if (n < N) {
#pragma unroll 4
    for (int k = ks; k < K; k += 8) {
      float x = __ldg(a + k);
      if (a_scale) x = fmaxf(fmaf(__ldg(a_scale + k), x, __ldg(a_shift + k)), 0.0f);
      acc = fmaf(x, __ldg(B + (size_t)k * ldb + n), acc);
    }
  }
  __shared__ float sh[8][32];
  sh[ks][threadIdx.x & 31] = acc;
  __syncthreads();
  if (ks == 0 && n < N) {
    for (int j = 1; j < 8; ++j) acc += sh[j][threadIdx.x];
    y0[n] = acc + (bias ? bias[n] : 0.0f);
  }
}

// (b) C[M,N] = A[M,K] . B^T with B stored [N,K] row-major (B(k,n) = B[n*ldb + k]), N <= 16, K % 4 == 0, K <= 1024:
//     one warp per row: lanes stride over K with 128-bit loads, N warp reductions.  (A 4-rows-per-warp / 8-lanes-per-row
//     mapping with 3 shuffle steps per output was measured slower: 426 against 305 us at N = 12.)
__global__ void __launch_bounds__(256) skinny_n_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B,
                                                      long long sbn, const float* __restrict__ bias, float* __restrict__ C,
                                                      long long ldc, int M, int N, int K, long long sbk = 1) {
  extern __shared__ float sB[];                      // [N][K]; B(k,n) = B[k*sbk + n*sbn]
  for (int i = threadIdx.x; i < N * K; i += 256) sB[i] = B[(long long)(i % K) * sbk + (long long)(i / K) * sbn];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long m = (long long)blockIdx.x * 8 + warp; m < M; m += (long long)gridDim.x * 8) {
    const float* a = A + m * lda;
    float acc[kSkinnyMax];
#pragma unroll
    for (int j = 0; j < kSkinnyMax; ++j) acc[j] = 0.0f;
    for (int k = lane * 4; k < K; k += 128) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(a + k));
#pragma unroll
      for (int j = 0; j < kSkinnyMax; ++j) {
        if (j < N) {
          const float4 w = *reinterpret_cast<const float4*>(&sB[j * K + k]);
          acc[j] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[j]))));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kSkinnyMax; ++j) {
      if (j < N) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
      }
    }
    if (lane < N) {
      float v = 0.0f;
#pragma unroll
      for (int j = 0; j < kSkinnyMax; ++j) if (j == lane) v = acc[j];
      C[m * ldc + lane] = v + (bias ? bias[lane] : 0.0f);
    }
  }
}

// (c) C[M,N] += A^T . B over K rows: A(m,k) = A[k*lda + m] (M <= 16 contiguous per row), B(k,n) = B[k*ldb + n], N % 4 == 0,
//     N <= 1024.  thread = 4 columns x all M; rows strided over (row lane, block); block-level reduction, then atomics
//     into a zero-initialised C.
__global__ void __launch_bounds__(256) skinny_m_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B,
                                                      long long ldb, float* __restrict__ C, long long ldc, int M, int N, int K) {
  const int n4 = N / 4, row_lanes = 256 / n4;
  const int cq = (threadIdx.x % n4) * 4, rl = threadIdx.x / n4;
  float4 acc[kSkinnyMax];
#pragma unroll
  for (int i = 0; i < kSkinnyMax; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < row_lanes) {
    for (long long k = (long long)blockIdx.x * row_lanes + rl; k < K; k += (long long)gridDim.x * row_lanes) {
      const float4 dy = __ldg(reinterpret_cast<const float4*>(B + k * ldb + cq));
      const float* a = A + k * lda;
#pragma unroll
      for (int i = 0; i < kSkinnyMax; ++i) {
        if (i < M) {
          const float x = __ldg(a + i);
          acc[i].x = fmaf(x, dy.x, acc[i].x); acc[i].y = fmaf(x, dy.y, acc[i].y);
          acc[i].z = fmaf(x, dy.z, acc[i].z); acc[i].w = fmaf(x, dy.w, acc[i].w);
        }
      }
    }
  }
  extern __shared__ float sC[];                      // [M][N] block partial sums
  for (int i = threadIdx.x; i < M * N; i += 256) sC[i] = 0.0f;
  __syncthreads();
  if (rl < row_lanes) {
#pragma unroll
    for (int i = 0; i < kSkinnyMax; ++i) {
      if (i < M) {
        atomicAdd(&sC[i * N + cq], acc[i].x); atomicAdd(&sC[i * N + cq + 1], acc[i].y);
        atomicAdd(&sC[i * N + cq + 2], acc[i].z); atomicAdd(&sC[i * N + cq + 3], acc[i].w);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < M * N; i += 256) atomicAdd(C + (long long)(i / N) * ldc + (i % N), sC[i]);
}

}  // namespace t3d
