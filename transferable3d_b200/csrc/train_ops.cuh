// fp32 CUDA-core kernels of the training step (BoxPC-Fit training, train_boxpc.py:219-300, and the building
// blocks of the semi-supervised step): general strided GEMM with split-K (forward / dgrad / wgrad of
// tf_util.conv2d / fully_connected), training-mode batch norm (tf.contrib.layers.batch_norm, tf_util.py:1645-1664)
// forward + backward fused with ReLU, max-pool over points with argmax + scatter backward, dropout, the
// BoxPC loss (boxpc_sunrgbd.py:106-193) forward + backward, and TF-style Adam on a flat arena.
#pragma once
#include "common.cuh"

namespace t3d {

// ----------------------------------------------------------------------------- general GEMM
// C[M,N] (+)= sum_k A(m,k) * B(k,n) with arbitrary element strides; split-K partials are atomically added
// into a zero-initialised C.  Either stride of A (and of B) must be 1.
struct GemmArgs {
  const float* A; long long sam, sak;
  const float* B; long long sbk, sbn;
  float* C; int ldc;
  int M, N, K, splitk;
  const float* bias;        // [N], added by split 0 (may be null)
  // fused batch-norm statistics of C (tensor-core path, splitk == 1): st_sum[n] += sum_m (C - st_shift[n]),
  // st_sq[n] += sum_m (C - st_shift[n])^2; zero-initialised by the caller.  null: off
  float* st_sum; float* st_sq; const float* st_shift;
  // fused max-pool of a forward-only lazy BN layer (C == null: the output is never stored): per group of pool_rows rows the
  // column maximum and minimum of the pre-BN output, as order-preserving keys (f32_ordered) merged by atomic max / min into
  // pool_max / pool_min [groups, N] (initialised to 0 / 0xffffffff).  relu(a x + b) is monotone in x, so the pooled value is
  // relu(a max + b) for a >= 0 and relu(a min + b) otherwise (pool_bn_finish_kernel).  pool_rows % 128 == 0.
  unsigned* pool_max; unsigned* pool_min; int pool_rows;
};

__global__ void __launch_bounds__(256) gemm_f32_kernel(const GemmArgs a) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kchunk = ((a.K + a.splitk - 1) / a.splitk + BK - 1) / BK * BK;
  const int kbeg = blockIdx.z * kchunk, kend = min(a.K, kbeg + kchunk);
  float acc[4][4] = {};
  const bool a_kmajor = (a.sak == 1), b_nmajor = (a.sbn == 1);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    if (a_kmajor) {
      const int r = tid >> 2, kk = (tid & 3) * 4, gm = m0 + r;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int gk = k0 + kk + e;
        As[kk + e][r] = (gm < a.M && gk < kend) ? a.A[(long long)gm * a.sam + gk] : 0.0f;
      }
    } else {
      const int kk = tid >> 4, r = (tid & 15) * 4, gk = k0 + kk;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int gm = m0 + r + e;
        As[kk][r + e] = (gm < a.M && gk < kend) ? a.A[(long long)gk * a.sak + gm] : 0.0f;
      }
    }
    if (b_nmajor) {
      const int kk = tid >> 4, nn = (tid & 15) * 4, gk = k0 + kk;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int gn = n0 + nn + e;
        Bs[kk][nn + e] = (gk < kend && gn < a.N) ? a.B[(long long)gk * a.sbk + gn] : 0.0f;
      }
    } else {
      const int nn = tid >> 2, kk = (tid & 3) * 4, gn = n0 + nn;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int gk = k0 + kk + e;
        Bs[kk + e][nn] = (gk < kend && gn < a.N) ? a.B[(long long)gn * a.sbn + gk] : 0.0f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= a.N) continue;
      float v = acc[i][j];
      if (a.bias && blockIdx.z == 0) v += a.bias[gn];
      if (a.splitk > 1) atomicAdd(a.C + (size_t)gm * a.ldc + gn, v);
      else a.C[(size_t)gm * a.ldc + gn] = v;
    }
  }
}

// activation ids follow include/t3d_b200.h: 0 none, 1 relu, 2 leaky_relu(0.2), 3 tanh
__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == 1) return fmaxf(v, 0.0f);
  if (act == 2) return v > 0.0f ? v : 0.2f * v;
  if (act == 3) return tanhf(v);
  return v;
}
// derivative expressed through the OUTPUT value (what the forward pass keeps)
__device__ __forceinline__ float act_grad_from_out(float out, int act) {
  if (act == 1) return out > 0.0f ? 1.0f : 0.0f;
  if (act == 2) return out > 0.0f ? 1.0f : 0.2f;
  if (act == 3) return 1.0f - out * out;
  return 1.0f;
}

// ----------------------------------------------------------------------------- column statistics
// out0[c] += sum_r f0, out1[c] += sum_r f1 over the rows of X[M,C] (zero-initialised outputs, atomics).
//  mode 0: f0 = x - shift[c],  f1 = (x - shift[c])^2          (BN forward: mean / variance; shift = row 0 of X keeps
//                                                            E[d^2] - E[d]^2 free of cancellation when |mean| >> std)
//  mode 1: f0 = dy,           f1 = dy * xhat                 (BN backward; dy = dOut * act'(out) if out != null)
//          xhat = (y - mean[c]) * rstd[c]
struct ColStatArgs {
  const float* X;           // mode 0: x ; mode 1: dOut
  const float* out;         // mode 1: post-activation output (ReLU mask) or null
  const float* y;           // mode 1: pre-BN values
  const float* mean; const float* rstd;
  float* o0; float* o1;
  int M, C, mode, act;
  // lazy BN layers keep no post-activation tensor: the ReLU mask is recomputed as a_scale[c] * y + a_shift[c] > 0, the
  // expression the consumer's loader evaluated in the forward pass (out must be null then)
  const float* a_scale; const float* a_shift;
};
__global__ void __launch_bounds__(256) colstats_kernel(const ColStatArgs a) {
  // block: 32 columns x 8 row-lanes; grid.x over column groups, grid.y over row chunks
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  const int rows_per_block = (a.M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(a.M, r0 + rows_per_block);
  float s0 = 0.f, s1 = 0.f;
  if (c < a.C) {
    const float mu = a.mode == 1 ? a.mean[c] : 0.f, rs = a.mode == 1 ? a.rstd[c] : 0.f;
    const float shift = (a.mode == 0 && a.y) ? a.y[c] : 0.f;
    const float asc = a.a_scale ? a.a_scale[c] : 0.f, ash = a.a_scale ? a.a_shift[c] : 0.f;
    // four rows in flight per thread (the loop carried 2 loads: 16 KB in flight per SM, 3.4 TB/s)
    int r = r0 + rl;
    if (a.mode == 0) {
      for (; r + 24 < r1; r += 32) {
        float x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = a.X[(size_t)(r + 8 * u) * a.C + c];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const float d = x[u] - shift; s0 += d; s1 = fmaf(d, d, s1); }
      }
      for (; r < r1; r += 8) { const float d = a.X[(size_t)r * a.C + c] - shift; s0 += d; s1 = fmaf(d, d, s1); }
    } else {
      auto term = [&](float dy, float yv, float ov) {
        if (a.out) dy *= act_grad_from_out(ov, a.act);
        else if (a.a_scale) dy = fmaf(asc, yv, ash) > 0.0f ? dy : 0.0f;
        s0 += dy; s1 = fmaf(dy, (yv - mu) * rs, s1);
      };
      for (; r + 24 < r1; r += 32) {
        float d[4], yv[4], ov[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const size_t i = (size_t)(r + 8 * u) * a.C + c;
          d[u] = a.X[i]; yv[u] = a.y[i]; ov[u] = a.out ? a.out[i] : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) term(d[u], yv[u], ov[u]);
      }
      for (; r < r1; r += 8) {
        const size_t i = (size_t)r * a.C + c;
        term(a.X[i], a.y[i], a.out ? a.out[i] : 0.0f);
      }
    }
  }
  __shared__ float sh0[8][32], sh1[8][32];
  sh0[rl][threadIdx.x & 31] = s0; sh1[rl][threadIdx.x & 31] = s1;
  __syncthreads();
  if (rl == 0 && c < a.C) {
    for (int k = 1; k < 8; ++k) { s0 += sh0[k][threadIdx.x]; s1 += sh1[k][threadIdx.x]; }
    atomicAdd(a.o0 + c, s0); atomicAdd(a.o1 + c, s1);
  }
}

// finalize BN statistics: mean, biased var -> rstd; moving <- decay*moving + (1-decay)*batch (unbiased var into
// the moving variance, TF1 fused-BN behaviour)
__global__ void bn_finalize_kernel(const float* sum, const float* sumsq, const float* shift, int M, int C, float eps, float decay,
                                   float* mean, float* rstd, float* moving_mean, float* moving_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float md = sum[c] / (float)M;                       // mean of (x - shift)
  const float var = fmaxf(sumsq[c] / (float)M - md * md, 0.0f);
  const float mu = md + (shift ? shift[c] : 0.0f);
  mean[c] = mu; rstd[c] = 1.0f / sqrtf(var + eps);
  if (moving_mean) {
    moving_mean[c] = decay * moving_mean[c] + (1.0f - decay) * mu;
    const float unb = var * ((float)M / (float)max(M - 1, 1));
    moving_var[c] = decay * moving_var[c] + (1.0f - decay) * unb;
  }
}

// same, plus the folded affine map of the lazy BN layers: scale = gamma * rstd, shift = beta - mean * scale
__global__ void bn_finalize_affine_kernel(const float* sum, const float* sumsq, const float* shift, int M, int C, float eps, float decay,
                                          const float* gamma, const float* beta, float* mean, float* rstd, float* a_scale,
                                          float* a_shift, float* moving_mean, float* moving_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float md = sum[c] / (float)M;
  const float var = fmaxf(sumsq[c] / (float)M - md * md, 0.0f);
  const float mu = md + (shift ? shift[c] : 0.0f);
  const float rs = 1.0f / sqrtf(var + eps);
  mean[c] = mu; rstd[c] = rs;
  const float sc = gamma[c] * rs;
  a_scale[c] = sc; a_shift[c] = beta[c] - mu * sc;
  if (moving_mean) {
    moving_mean[c] = decay * moving_mean[c] + (1.0f - decay) * mu;
    const float unb = var * ((float)M / (float)max(M - 1, 1));
    moving_var[c] = decay * moving_var[c] + (1.0f - decay) * unb;
  }
}

// out = act(gamma * (y - mean) * rstd + beta)
__global__ void bn_apply_kernel(const float* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ rstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out,
                                size_t total, int C, int act) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  float v = gamma[c] * (y[i] - mean[c]) * rstd[c] + beta[c];
  out[i] = act_apply(v, act);
}

// dY = gamma*rstd * (dy - s1/M - xhat*s2/M), dy = dOut*act'(out) ; written over dOut (in place)
__global__ void bn_backward_kernel(float* __restrict__ dOut, const float* __restrict__ out, const float* __restrict__ y,
                                   const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                                   const float* __restrict__ s1, const float* __restrict__ s2, size_t total, int C, int M, int act,
                                   const float* __restrict__ a_scale = nullptr, const float* __restrict__ a_shift = nullptr) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  float dy = dOut[i];
  if (out) dy *= act_grad_from_out(out[i], act);
  else if (a_scale) dy = fmaf(a_scale[c], y[i], a_shift[c]) > 0.0f ? dy : 0.0f;
  const float xh = (y[i] - mean[c]) * rstd[c];
  const float inv = 1.0f / (float)M;
  dOut[i] = gamma[c] * rstd[c] * (dy - s1[c] * inv - xh * s2[c] * inv);
}

// ----------------------------------------------------------------------------- max-pool over points with argmax
// rowmask != null: the pooled tensor is x * rowmask[row] (the `net * mask` in front of the max-pool of the masked stacks,
// semisup_models.py:184-185, 240-241) without materialising the product; its backward scales the routed gradient by the
// mask of the arg-max row -- identical to multiplying afterwards, two passes over the activation fewer each way.
// a_scale != null: x is the pre-BN tensor of a lazy BN layer and the pooled value is relu(a_scale[c] * x + a_shift[c]).
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ rowmask, int B, int N, int C,
                                   float* __restrict__ out, int* __restrict__ arg, const float* __restrict__ a_scale = nullptr,
                                   const float* __restrict__ a_shift = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;        // over B*C
  if (i >= B * C) return;
  const int b = i / C, c = i % C;
  const float* p = x + (size_t)b * N * C + c;
  const float* rm = rowmask ? rowmask + (size_t)b * N : nullptr;
  const bool lazy = a_scale != nullptr;
  const float sc = lazy ? a_scale[c] : 1.0f, sh = lazy ? a_shift[c] : 0.0f;
  float m = lazy ? fmaxf(fmaf(sc, p[0], sh), 0.0f) : p[0];
  if (rm) m *= rm[0];
  int am = 0;
  int n = 1;
  for (; n + 8 <= N; n += 8) {                                  // 8 rows in flight; compared in order: first max wins
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = p[(size_t)(n + j) * C];
    if (lazy) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(sc, v[j], sh), 0.0f);
    }
    if (rm) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= rm[n + j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) if (v[j] > m) { m = v[j]; am = n + j; }
  }
  for (; n < N; ++n) {
    float v = p[(size_t)n * C];
    if (lazy) v = fmaxf(fmaf(sc, v, sh), 0.0f);
    if (rm) v *= rm[n];
    if (v > m) { m = v; am = n; }
  }
  out[i] = m; arg[i] = am;
}

// Row-split variant: blockIdx.z owns a chunk of the N rows, so that enough loads are in flight to fill HBM when B * C / 256
// blocks are too few (3.5 blocks per SM at B = 256, C = 512: 2.9 TB/s).  Chunks are merged by a 64-bit atomic max on
// (order-preserving image of the value) << 32 | ~row: the largest value wins, the smallest row among equal values -- the
// "first max wins" of the serial kernel.  maxpool_split_finish_kernel unpacks into (out, arg).
__device__ __forceinline__ uint32_t f32_ordered(float f) {
  const uint32_t u = __float_as_uint(f + 0.0f);                 // -0 -> +0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_unordered(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
__global__ void __launch_bounds__(256) maxpool_split_kernel(const float* __restrict__ x, const float* __restrict__ rowmask, int B, int N,
                                                           int C, int rows_per_chunk, unsigned long long* __restrict__ keys,
                                                           const float* __restrict__ a_scale, const float* __restrict__ a_shift) {
  const int c = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  if (c >= C) return;
  const int n0 = blockIdx.z * rows_per_chunk, n1 = min(N, n0 + rows_per_chunk);
  if (n0 >= n1) return;
  const float* p = x + (size_t)b * N * C + c;
  const float* rm = rowmask ? rowmask + (size_t)b * N : nullptr;
  const bool lazy = a_scale != nullptr;
  const float sc = lazy ? a_scale[c] : 1.0f, sh = lazy ? a_shift[c] : 0.0f;
  auto val = [&](float v, int n) {
    if (lazy) v = fmaxf(fmaf(sc, v, sh), 0.0f);
    if (rm) v *= rm[n];
    return v;
  };
  float m = val(p[(size_t)n0 * C], n0); int am = n0;
  int n = n0 + 1;
  for (; n + 8 <= n1; n += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = p[(size_t)(n + j) * C];
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] = val(v[j], n + j); if (v[j] > m) { m = v[j]; am = n + j; } }
  }
  for (; n < n1; ++n) { const float v = val(p[(size_t)n * C], n); if (v > m) { m = v; am = n; } }
  const unsigned long long key = ((unsigned long long)f32_ordered(m) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)am);
  atomicMax(keys + (size_t)b * C + c, key);
}
__global__ void maxpool_split_finish_kernel(const unsigned long long* __restrict__ keys, int total, float* __restrict__ out,
                                            int* __restrict__ arg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const unsigned long long k = keys[i];
  out[i] = f32_unordered((uint32_t)(k >> 32));
  arg[i] = (int)(0xffffffffu - (uint32_t)(k & 0xffffffffull));
}

// pooled[g, c] = relu(a_scale[c] * (a_scale[c] >= 0 ? max : min) + a_shift[c]) from the keys of the fused max-pool epilogue
__global__ void pool_bn_finish_kernel(const unsigned* __restrict__ kmax, const unsigned* __restrict__ kmin, const float* __restrict__ a_scale,
                                      const float* __restrict__ a_shift, int total, int C, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = i % C;
  const float sc = a_scale[c];
  const float x = sc >= 0.0f ? f32_unordered(kmax[i]) : f32_unordered(kmin[i]);
  out[i] = fmaxf(fmaf(sc, x, a_shift[c]), 0.0f);
}

// ---- backward of [BN -> ReLU -> (row mask) -> max-pool] without the dense pooled-gradient tensor ------------------------
// The gradient w.r.t. the BN output is non-zero only at the B x C arg-max elements, so the two batch-norm reductions
// (s1 = sum dy, s2 = sum dy * xhat) are O(B C) gathers instead of two passes over the B N x C tensor, and the BN input
// gradient  dY = gamma rstd (dy - s1 / M - xhat s2 / M)  is one dense pass that reads y and writes dY plus a B x C
// scatter.  Replaces maxpool_bwd (memset + scatter) + colstats mode 1 + bn_backward: 8 passes over the widest
// activation of every stack become 2.
struct PoolBnBwdArgs {
  const float* g;            // [B, C] gradient w.r.t. the pooled features
  const int* arg;            // [B, C] arg-max rows
  const float* rowmask;      // [B * N] or null
  const float* y;            // [B * N, C] pre-BN values
  const float* mean; const float* rstd; const float* gamma;
  const float* a_scale; const float* a_shift;      // ReLU mask = a_scale * y + a_shift > 0
  float* s1; float* s2;      // [C] zero-initialised (= d beta, d gamma)
  float* dY;                 // [B * N, C]
  int B, N, C;
};
__device__ __forceinline__ float pool_bn_dy(const PoolBnBwdArgs& a, int i, size_t& at) {
  const int b = i / a.C, c = i % a.C;
  const size_t row = (size_t)b * a.N + a.arg[i];
  at = row * a.C + c;
  const float yv = a.y[at];
  float dy = fmaf(a.a_scale[c], yv, a.a_shift[c]) > 0.0f ? a.g[i] : 0.0f;
  if (a.rowmask) dy *= a.rowmask[row];
  return dy;
}
__global__ void __launch_bounds__(256) pool_bn_stats_kernel(const PoolBnBwdArgs a) {      // grid (C / 32.., B chunks), block (32, 8)
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s1 = 0.f, s2 = 0.f;
  if (c < a.C) {
    const float mu = a.mean[c], rs = a.rstd[c];
    for (int b = blockIdx.y * 8 + threadIdx.y; b < a.B; b += gridDim.y * 8) {
      size_t at;
      const float dy = pool_bn_dy(a, b * a.C + c, at);
      s1 += dy; s2 = fmaf(dy, (a.y[at] - mu) * rs, s2);
    }
  }
  __shared__ float sh[2][8][32];
  sh[0][threadIdx.y][threadIdx.x] = s1; sh[1][threadIdx.y][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y == 0 && c < a.C) {
    for (int k = 1; k < 8; ++k) { s1 += sh[0][k][threadIdx.x]; s2 += sh[1][k][threadIdx.x]; }
    atomicAdd(a.s1 + c, s1); atomicAdd(a.s2 + c, s2);
  }
}
// dense part: dY = -gamma rstd (s1 + xhat s2) / M     (float4 over [M, C], parameters staged per block)
__global__ void __launch_bounds__(256) pool_bn_dense_kernel(const float4* __restrict__ y, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ s1, const float* __restrict__ s2,
                                                           float4* __restrict__ dY, unsigned n4, unsigned C4, int M) {
  extern __shared__ float4 pb_sp[];                  // [3][C4]: mean, k1 = -gamma rstd s1 / M, k2 = -gamma rstd^2 s2 / M
  const float inv = 1.0f / (float)M;
  for (unsigned i = threadIdx.x; i < C4; i += 256) {
    float m[4], k1[4], k2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const unsigned c = 4u * i + e;
      const float gr = gamma[c] * rstd[c];
      m[e] = mean[c]; k1[e] = -gr * s1[c] * inv; k2[e] = -gr * rstd[c] * s2[c] * inv;
    }
    pb_sp[i] = make_float4(m[0], m[1], m[2], m[3]); pb_sp[C4 + i] = make_float4(k1[0], k1[1], k1[2], k1[3]);
    pb_sp[2 * C4 + i] = make_float4(k2[0], k2[1], k2[2], k2[3]);
  }
  __syncthreads();
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    const unsigned base = (blockIdx.x * 4 + it) * 1024u + threadIdx.x;
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) if (base + u * 256u < n4) v[u] = y[base + u * 256u];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned q = base + u * 256u;
      if (q >= n4) continue;
      const unsigned c = q % C4;
      const float4 m = pb_sp[c], k1 = pb_sp[C4 + c], k2 = pb_sp[2 * C4 + c];
      float4 o;
      o.x = fmaf(v[u].x - m.x, k2.x, k1.x); o.y = fmaf(v[u].y - m.y, k2.y, k1.y);
      o.z = fmaf(v[u].z - m.z, k2.z, k1.z); o.w = fmaf(v[u].w - m.w, k2.w, k1.w);
      dY[q] = o;
    }
  }
}
__global__ void __launch_bounds__(256) pool_bn_dense_scalar_kernel(const float* __restrict__ y, const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                  const float* __restrict__ s1, const float* __restrict__ s2,
                                                                  float* __restrict__ dY, size_t total, int C, int M) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const float inv = 1.0f / (float)M, gr = gamma[c] * rstd[c];
  dY[i] = fmaf(y[i] - mean[c], -gr * rstd[c] * s2[c] * inv, -gr * s1[c] * inv);
}
// sparse part (after the dense pass): dY[arg-max element] += gamma rstd dy; an element can be the arg-max of one (b, c) only
__global__ void __launch_bounds__(256) pool_bn_scatter_kernel(const PoolBnBwdArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B * a.C) return;
  size_t at;
  const float dy = pool_bn_dy(a, i, at);
  const int c = i % a.C;
  a.dY[at] += a.gamma[c] * a.rstd[c] * dy;
}
__global__ void maxpool_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ arg, const float* __restrict__ rowmask,
                                   int B, int N, int C, float* __restrict__ dx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i % C;
  const size_t row = (size_t)b * N + arg[i];
  dx[row * C + c] = rowmask ? dout[i] * rowmask[row] : dout[i];       // dx is zero-initialised by the caller
}

// ---- row-sparse backward through a frozen stack behind a max-pool ----------------------------------------------------------
// The gradient entering a max-pool's input is non-zero only in the rows that are the arg-max of some channel: at most
// min(C, N) of the N rows of a frustum.  Through layers WITHOUT batch statistics (eval-mode BN folded: the frozen BoxPC
// branch of train_semisup_adv.py:364-388) the input gradient stays confined to those rows, so the whole backward chain
// runs on the compacted rows.  pool_rows_kernel: per frustum, the ascending list of distinct arg-max rows (padded with -1
// to S slots) and, per channel, the slot of its arg-max row.
__global__ void __launch_bounds__(256) pool_rows_kernel(const int* __restrict__ arg, int N, int C, int S, int* __restrict__ rows,
                                                       int* __restrict__ slot, int* __restrict__ count) {
  extern __shared__ unsigned pr_bits[];               // [words] bitmap, then [words] exclusive popcount prefix
  const int b = blockIdx.x, words = (N + 31) / 32;
  unsigned* pre = pr_bits + words;
  for (int w = threadIdx.x; w < words; w += 256) pr_bits[w] = 0u;
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) { const int r = arg[(size_t)b * C + c]; atomicOr(&pr_bits[r >> 5], 1u << (r & 31)); }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned run = 0;
    for (int w = 0; w < words; ++w) { pre[w] = run; run += __popc(pr_bits[w]); }
    count[b] = (int)run;
  }
  __syncthreads();
  const int cnt = min(count[b], S);
  for (int c = threadIdx.x; c < C; c += 256) {
    const int r = arg[(size_t)b * C + c];
    const int sl = (int)(pre[r >> 5] + __popc(pr_bits[r >> 5] & ((1u << (r & 31)) - 1u)));
    slot[(size_t)b * C + c] = sl;
    if (sl < S) rows[(size_t)b * S + sl] = r;
  }
  for (int s2 = cnt + threadIdx.x; s2 < S; s2 += 256) rows[(size_t)b * S + s2] = -1;
}
// dst[b*S + s, :] = src[b*N + rows[b,s], :]  (zeros where rows = -1); C % 4 == 0 takes 128-bit copies
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ rows, int N, int S, int C,
                                                         size_t total_rows, float* __restrict__ dst) {
  const int per_row = (C % 4 == 0) ? C / 4 : C;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_rows * per_row) return;
  const size_t row = i / per_row;
  const int j = (int)(i % per_row);
  const int b = (int)(row / S), r = rows[row];
  if (C % 4 == 0) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= 0) v = *reinterpret_cast<const float4*>(src + ((size_t)b * N + r) * C + 4 * j);
    *reinterpret_cast<float4*>(dst + row * C + 4 * j) = v;
  } else {
    dst[row * C + j] = r >= 0 ? src[((size_t)b * N + r) * C + j] : 0.0f;
  }
}
// dst[b*S + slot[b,c], c] = g[b,c]   (dst zero-initialised by the caller)
__global__ void __launch_bounds__(256) scatter_pool_grad_kernel(const float* __restrict__ g, const int* __restrict__ slot, int B, int C, int S,
                                                               float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i % C, sl = slot[i];
  if (sl < S) dst[((size_t)b * S + sl) * C + c] = g[i];
}

// out = x * mask * scale  (tf.nn.dropout forward and backward with the same keep mask, scale = 1/keep_prob)
__global__ void scale_mask_kernel(const float* __restrict__ x, const float* __restrict__ mask, float scale, float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[i] * mask[i] * scale;
}

// ----------------------------------------------------------------------------- BoxPC loss (boxpc_sunrgbd.py:106-193)
// per sample: cls = softmaxCE(onehot(iou > fit_bound), fit logits); delta = wc*mean3 huber(dc) + ws*mean3 huber(ds) + wa*huber(da)
// total = mean_B(w_cls*cls + w_delta*delta).  out9 = [dc(3), ds(3), da, l0, l1].  grad (B,9) = d total / d out9.
struct BoxpcLossArgs {
  const float* out9; const float* y_iou; const float* y_dc; const float* y_ds; const float* y_da;
  int B; float fit_bound, w_cls, w_delta, wc, ws, wa; int huber;     // huber=1, else mse
  float* cls_losses; float* delta_losses; float* total; float* grad;
  // class-confidence weighting (boxpc_sunrgbd.py:76-91, 166-177), p1 = softmax(fit logits)[1]:
  //   pred_weigh  1: the deltas entering the loss are out9[0:7] * (1 - p1)      (BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF)
  //   loss_weigh  1: delta loss * (1 - p1) (..._LOSS_BY_CLS_CONF), 2: * (1 - y_iou) (..._LOSS_BY_CLS_GT)
  //   stop_grad   1: p1 is a constant in both (BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA)
  int pred_weigh, loss_weigh, stop_grad;
};
__device__ __forceinline__ void huber1(float err, float& loss, float& dloss) {   // tf.losses.huber_loss delta=1 on (pred-label)
  const float a = fabsf(err), q = fminf(a, 1.0f);
  loss = 0.5f * q * q + (a - q);
  dloss = a <= 1.0f ? err : (err > 0.f ? 1.0f : -1.0f);
}
__global__ void boxpc_loss_kernel(const BoxpcLossArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float contrib = 0.f;
  if (b < a.B) {
    const float* o = a.out9 + (size_t)b * 9;
    float g[9];
    const float l0 = o[7], l1 = o[8], mx = fmaxf(l0, l1);
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx), z = e0 + e1;
    const float p0 = e0 / z, p1 = e1 / z;
    const int lab = a.y_iou[b] > a.fit_bound ? 1 : 0;
    const float cls = logf(z) + mx - (lab ? l1 : l0);
    const float inv_b = 1.0f / (float)a.B;
    g[7] = a.w_cls * inv_b * (p0 - (lab == 0 ? 1.f : 0.f));
    g[8] = a.w_cls * inv_b * (p1 - (lab == 1 ? 1.f : 0.f));
    const float wp = a.pred_weigh ? 1.0f - p1 : 1.0f;                                     // scales the predicted deltas
    const float wl = a.loss_weigh == 1 ? 1.0f - p1 : (a.loss_weigh == 2 ? 1.0f - a.y_iou[b] : 1.0f);      // scales the loss
    float lc = 0.f, ls = 0.f, la, d;
    float dot = 0.f;                       // sum_k dD/dd'_k * o_k : how D moves with wp
    for (int k = 0; k < 3; ++k) {
      float l;
      const float ec = o[k] * wp - a.y_dc[b * 3 + k], es = o[3 + k] * wp - a.y_ds[b * 3 + k];
      if (a.huber) huber1(ec, l, d); else { l = ec * ec; d = 2.f * ec; }
      lc += l; d *= a.wc / 3.0f; dot = fmaf(d, o[k], dot); g[k] = a.w_delta * inv_b * wl * wp * d;
      if (a.huber) huber1(es, l, d); else { l = es * es; d = 2.f * es; }
      ls += l; d *= a.ws / 3.0f; dot = fmaf(d, o[3 + k], dot); g[3 + k] = a.w_delta * inv_b * wl * wp * d;
    }
    const float ea = o[6] * wp - a.y_da[b];
    if (a.huber) huber1(ea, la, d); else { la = ea * ea; d = 2.f * ea; }
    d *= a.wa; dot = fmaf(d, o[6], dot); g[6] = a.w_delta * inv_b * wl * wp * d;
    const float D = a.wc * lc / 3.0f + a.ws * ls / 3.0f + a.wa * la;
    const float delta = wl * D;
    if (!a.stop_grad && (a.pred_weigh || a.loss_weigh == 1)) {
      // d total / d p1, then through the softmax: d p1 / d l1 = p0 p1 = - d p1 / d l0
      const float dp1 = a.w_delta * inv_b * ((a.loss_weigh == 1 ? -D : 0.0f) + (a.pred_weigh ? -wl * dot : 0.0f));
      g[8] += dp1 * p0 * p1;
      g[7] -= dp1 * p0 * p1;
    }
    if (a.cls_losses) a.cls_losses[b] = cls;
    if (a.delta_losses) a.delta_losses[b] = delta;
    if (a.grad) for (int k = 0; k < 9; ++k) a.grad[(size_t)b * 9 + k] = g[k];
    contrib = (a.w_cls * cls + a.w_delta * delta) * inv_b;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(a.total, contrib);          // total is zero-initialised by the caller
}

// ----------------------------------------------------------------------------- TF Adam
// lr_t = lr*sqrt(1-b2^t)/(1-b1^t) is computed by the host; theta -= lr_t * m / (sqrt(v) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n, float lr_t, float b1, float b2, float eps, float gscale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * gscale;
  const float mi = b1 * m[i] + (1.0f - b1) * gi;
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  p[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

}  // namespace t3d
