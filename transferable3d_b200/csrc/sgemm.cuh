// fp32 CUDA-core GEMM core shared by t3d_linear_f32 (fp32-mode layers, FC heads, the per-frustum conv6 bias) and
// t3d_gemm_f32 (training-step forward / dgrad / wgrad): C[M,N] = sum_k A(m,k) B(k,n) with element strides, one of each
// operand's strides being 1.
//
// 128 x 128 x 16 tiles, 256 threads, 8 x 8 register micro-tile per thread (two 4-wide halves 64 apart, so every
// shared-memory read is a conflict-free LDS.128), double-buffered shared memory with the next tile's global loads in
// flight while the current one is multiplied, 128-bit global loads when the operand is aligned.  Small problems
// (M < 128 or N < 96) stay on the 64 x 64 kernels of simt_ops.cuh / train_ops.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace t3d {

constexpr int kSgBM = 128, kSgBN = 128, kSgBK = 16;

struct SgemmOperands {
  const float* A; long long sam, sak;     // A(m,k) = A[m*sam + k*sak]
  const float* B; long long sbk, sbn;     // B(k,n) = B[k*sbk + n*sbn]
  int M, N, K;
  int kchunk;                             // K range of blockIdx.z: [z*kchunk, min(K, (z+1)*kchunk))
  int vecA, vecB;                         // 128-bit loads allowed (base and leading stride 16-byte aligned)
};

// loads one 128(rows) x 16(k) operand tile into registers: r[8] per thread.
//   unit-k operand (stride along k == 1): thread -> row = tid & 127, k segment = (tid >> 7) * 8, 8 consecutive k
//   unit-row operand (stride along rows == 1): thread -> k = tid >> 4, row segment = (tid & 15) * 8, 8 consecutive rows
template <bool UNIT_K>
__device__ __forceinline__ void sg_load_tile(const float* __restrict__ P, long long srow, long long sk, int row0, int nrows, int k0,
                                             int kend, bool vec, int tid, float (&r)[8]) {
  if (UNIT_K) {
    const int row = row0 + (tid & 127), k = k0 + (tid >> 7) * 8;
    if (row < nrows && vec && k + 8 <= kend) {
      const float4* q = reinterpret_cast<const float4*>(P + (long long)row * srow + k);
      const float4 v0 = q[0], v1 = q[1];
      r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w; r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] = (row < nrows && k + e < kend) ? P[(long long)row * srow + k + e] : 0.0f;
    }
  } else {
    const int k = k0 + (tid >> 4), row = row0 + (tid & 15) * 8;
    if (k < kend && vec && row + 8 <= nrows) {
      const float4* q = reinterpret_cast<const float4*>(P + (long long)k * sk + row);
      const float4 v0 = q[0], v1 = q[1];
      r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w; r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] = (k < kend && row + e < nrows) ? P[(long long)k * sk + row + e] : 0.0f;
    }
  }
}
// registers -> shared tile S[16][128]
template <bool UNIT_K>
__device__ __forceinline__ void sg_store_tile(float (*S)[kSgBM], int tid, const float (&r)[8]) {
  if (UNIT_K) {
    const int row = tid & 127, k = (tid >> 7) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) S[k + e][row] = r[e];           // a warp writes 32 consecutive floats of one k row
  } else {
    const int k = tid >> 4, row = (tid & 15) * 8;
    *reinterpret_cast<float4*>(&S[k][row]) = make_float4(r[0], r[1], r[2], r[3]);
    *reinterpret_cast<float4*>(&S[k][row + 4]) = make_float4(r[4], r[5], r[6], r[7]);
  }
}

// Blackwell's packed fp32 FMA (FFMA2): two independent round-to-nearest FMAs per instruction -- the same results as two
// scalar fmaf, at twice the rate of the 3-register FFMA (which issues every other cycle per scheduler on sm_100).
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a, float b0, float b1) {
  unsigned long long d, av, bv;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bv) : "f"(b0), "f"(b1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(av), "l"(bv));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

// acc[i][j]: rows m0 + (i < 4 ? ty*4 + i : 64 + ty*4 + i - 4), columns n0 + (j < 4 ? tx*4 + j : 64 + tx*4 + j - 4)
template <bool A_UNIT_K, bool B_UNIT_K>
__device__ __forceinline__ void sgemm128_mainloop(const SgemmOperands& a, float (&acc)[8][8]) {
  __shared__ __align__(16) float As[2][kSgBK][kSgBM];
  __shared__ __align__(16) float Bs[2][kSgBK][kSgBN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * kSgBM, n0 = blockIdx.y * kSgBN;
  const int kbeg = blockIdx.z * a.kchunk, kend = min(a.K, kbeg + a.kchunk);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  float ra[8], rb[8];
  sg_load_tile<A_UNIT_K>(a.A, a.sam, a.sak, m0, a.M, kbeg, kend, a.vecA != 0, tid, ra);
  sg_load_tile<B_UNIT_K>(a.B, a.sbn, a.sbk, n0, a.N, kbeg, kend, a.vecB != 0, tid, rb);
  sg_store_tile<A_UNIT_K>(As[0], tid, ra);
  sg_store_tile<B_UNIT_K>(Bs[0], tid, rb);
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += kSgBK) {
    const bool more = k0 + kSgBK < kend;
    if (more) {
      sg_load_tile<A_UNIT_K>(a.A, a.sam, a.sak, m0, a.M, k0 + kSgBK, kend, a.vecA != 0, tid, ra);
      sg_load_tile<B_UNIT_K>(a.B, a.sbn, a.sbk, n0, a.N, k0 + kSgBK, kend, a.vecB != 0, tid, rb);
    }
#pragma unroll
    for (int k = 0; k < kSgBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; j += 2) ffma2(acc[i][j], acc[i][j + 1], av[i], bv[j], bv[j + 1]);
    }
    if (more) {
      sg_store_tile<A_UNIT_K>(As[buf ^ 1], tid, ra);
      sg_store_tile<B_UNIT_K>(Bs[buf ^ 1], tid, rb);
      __syncthreads();
      buf ^= 1;
    }
  }
}

__device__ __forceinline__ int sg_row(int m0, int ty, int i) { return m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4); }
__device__ __forceinline__ int sg_col(int n0, int tx, int j) { return n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4); }

// ---- t3d_linear_f32 on the 128-tile core: Y = act(X.W + bias + gbias[row / rows_per_group]) * rowmask, optional max over
// the rows of each group (same contract as linear_f32_kernel, simt_ops.cuh).  Declared after LinearArgs / apply_act.
#ifdef T3D_SGEMM_WITH_EPILOGUES
__global__ void __launch_bounds__(256, 2) linear128_kernel(const LinearArgs a, const SgemmOperands o) {
  __shared__ float red[16][kSgBN];
  float acc[8][8];
  sgemm128_mainloop<true, false>(o, acc);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * kSgBM, n0 = blockIdx.y * kSgBN;
  const bool one_group = a.gmax && (a.rows_per_group % kSgBM == 0) && (m0 + kSgBM <= a.M);
  float cmax[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = sg_row(m0, ty, i);
    if (gm >= a.M) continue;
    const int g = a.rows_per_group > 0 ? gm / a.rows_per_group : 0;
    const float rm = a.rowmask ? a.rowmask[gm] : 1.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = sg_col(n0, tx, j);
      if (gn >= a.N) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[gn];
      if (a.gbias) v += a.gbias[(size_t)g * a.N + gn];
      v = apply_act(v, a.act) * rm;
      acc[i][j] = v;
      if (a.gmax) {
        if (one_group) cmax[j] = fmaxf(cmax[j], v);
        else atomic_max_f32(a.gmax + (size_t)g * a.N + gn, v);
      }
    }
    if (a.Y) {      // 128-bit stores where the row segment is whole and aligned
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int gn = sg_col(n0, tx, h * 4);
        float* y = a.Y + (size_t)gm * a.ldy + gn;
        if (gn + 4 <= a.N && (((uintptr_t)y) & 15) == 0) {
          *reinterpret_cast<float4*>(y) = make_float4(acc[i][h * 4], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (gn + j < a.N) y[j] = acc[i][h * 4 + j];
        }
      }
    }
  }
  if (one_group) {   // values are >= 0 here (ReLU / mask): reduce the tile's 128 rows before the atomics
#pragma unroll
    for (int j = 0; j < 8; ++j) red[ty][j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4] = cmax[j];
    __syncthreads();
    if (tid < kSgBN) {
      float m = red[0][tid];
#pragma unroll
      for (int r = 1; r < 16; ++r) m = fmaxf(m, red[r][tid]);
      const int gn = n0 + tid;
      if (gn < a.N) atomic_max_f32(a.gmax + (size_t)(m0 / a.rows_per_group) * a.N + gn, m);
    }
  }
}

// ---- t3d_gemm_f32 on the 128-tile core (same contract as gemm_f32_kernel, train_ops.cuh)
template <bool A_UNIT_K, bool B_UNIT_K>
__global__ void __launch_bounds__(256, 2) gemm128_kernel(const GemmArgs a, const SgemmOperands o) {
  float acc[8][8];
  sgemm128_mainloop<A_UNIT_K, B_UNIT_K>(o, acc);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * kSgBM, n0 = blockIdx.y * kSgBN;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = sg_row(m0, ty, i);
    if (gm >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = sg_col(n0, tx, j);
      if (gn >= a.N) continue;
      float v = acc[i][j];
      if (a.bias && blockIdx.z == 0) v += a.bias[gn];
      if (a.splitk > 1) atomicAdd(a.C + (size_t)gm * a.ldc + gn, v);
      else a.C[(size_t)gm * a.ldc + gn] = v;
    }
  }
}
#endif

inline bool sg_aligned16(const void* p, long long lead_stride) {
  return (((uintptr_t)p & 15) == 0) && (lead_stride % 4 == 0);
}

}  // namespace t3d
