// fp32-accurate GEMM on the tcgen05 tensor cores ("bf16 x 3"): the core behind t3d_linear_f32 (fp32-mode layers) and
// t3d_gemm_f32 (training-step forward / dgrad / wgrad) for every problem large enough to fill 128 x 128 tiles.
//
//   C[M,N] = sum_k A(m,k) B(k,n),  fp32 operands in HBM with element strides (one stride of each operand is 1).
//
// Each fp32 operand value x is split exactly into three bf16 pieces x = x1 + x2 + x3 (8 + 8 + 8 significant bits, by
// truncation: x1 = top 16 bits of x, x2 = top 16 bits of x - x1, x3 = x - x1 - x2; every subtraction is exact), and the
// product is evaluated as the six partial products whose weight is >= 2^-16 of the leading one:
//     a1 b1                                 -> TMEM accumulator "main"
//     a1 b2 + a2 b1 + a1 b3 + a2 b2 + a3 b1 -> TMEM accumulator "small" (2^-8 of main, so its rounding is invisible)
// The dropped terms (a2 b3, a3 b2, a3 b3) are <= 2^-24 relative, i.e. below fp32 round-off; bf16 x bf16 products are
// exact in the fp32 accumulator.  Cost: 6 tensor-core passes per fp32 FMA = 1/6 of the bf16 rate (~230 TFLOP/s of
// fp32-equivalent work against ~35 TFLOP/s for the CUDA-core SGEMM of sgemm.cuh) at fp32-level accuracy.
//
// Structure (one 128 x 128 output tile per CTA, 2 CTAs per SM so one CTA's epilogue hides under the other's main loop):
//   warps 0-3  load the A tile (128 rows x 32 k per stage), split it in registers, store the three bf16 images
//   warps 4-7  same for the B tile (128 output columns x 32 k)
//   warp  8    TMEM allocation + MMA issue (one elected lane): 2 k-steps x 6 tcgen05.mma per stage
//   warps 0-7  epilogue after the main loop: tcgen05.ld of both accumulators, main + small, bias / activation / mask /
//              group max / split-K atomics, 128-byte row segments straight to global memory
// Operand images are K-major SWIZZLE_128B blocks [128 rows x 64 k] (the layout of common.cuh's descriptors); a stage is
// one half (32 k = 16-byte chunks 4s..4s+3 of every row) of the block, so the two-stage ring costs no extra shared memory:
// 6 images x 16 KB = 96 KB per CTA.
#pragma once
#include "common.cuh"

namespace t3d {

constexpr int kXgBM = 128, kXgBN = 128, kXgBK = 32;
constexpr int kXgThreads = 288;
#ifndef XG_SETS
#define XG_SETS 2
#endif
#ifndef XG_MINB
#define XG_MINB 2
#endif
#ifndef XG_PRE_SETS
#define XG_PRE_SETS 3
#endif
constexpr int kXgPreSets = XG_PRE_SETS;                        // register sets (16 values each) of the pre-split-B loaders
#ifndef XG_DBG
#define XG_DBG 0      // measurement only: 1 = no global loads (A), 2 = no image stores, 4 = no MMAs, 8 = no B copy, 16 = no epilogue
#endif
constexpr int kXgSets = XG_SETS;                               // register sets of the loaders (stages in flight + 1)
constexpr int kXgMinBlocks = XG_MINB;                          // CTAs per SM the register budget is compiled for
constexpr int kXgMaxKChunk = 2048;                            // longest K range accumulated in TMEM by one CTA
constexpr uint32_t kXgImage = 128 * 128;                       // bytes of one [128 x 64] bf16 image
constexpr uint32_t kXgBars = 6 * kXgImage;                     // barrier block offset
constexpr int kXgLazyMaxK = 512;                               // lazy BN of a k-contiguous A: scale / shift tables in shared memory
constexpr uint32_t kXgParams = kXgBars + 64;                   // [a_scale K][a_shift K] fp32 (one-tile kernels: K <= kXgLazyMaxK)
constexpr uint32_t kXgSmemBytes = 6 * kXgImage + 64 + 8 * kXgLazyMaxK + 1024;    // + barriers / TMEM slot + tables + 1024-byte alignment slack

struct XgOperands {
  const float* A; long long lda;      // UNIT_K: A(m,k) = A[m*lda + k];  else A(m,k) = A[k*lda + m]
  const float* B; long long ldb;      // UNIT_K: B(k,n) = B[n*ldb + k];  else B(k,n) = B[k*ldb + n]
  int M, N, K;
  int kchunk;                         // K range of blockIdx.y: [y*kchunk, min(K, (y+1)*kchunk)), multiple of 32
  int vecA, vecB;                     // 128-bit loads allowed for a UNIT_K operand (16-byte aligned base and ld % 4 == 0)
  int ntn;                            // number of column tiles; blockIdx.x = m_tile * ntn + n_tile
  unsigned long long* trace;          // debug: clock64 timeline of the middle CTA (t3d_set_trace_buffer) or null
  const uint8_t* bpre;                // B_PRE kernels: pre-split images of B, [n_tile][k_block of 64][part][16 KB], else null
  int nkb;                            // k blocks per n tile in bpre
  // Lazy batch norm of the A operand (training forward / wgrad): the value multiplied is relu(a_scale[c] * A + a_shift[c])
  // with c the CHANNEL of the stored pre-BN activation (its k index when A is k-contiguous, its row m otherwise), so the
  // post-BN activation of the previous layer never exists in HBM.  null: A is used as stored.  K % 32 == 0 required.
  const float* a_scale; const float* a_shift;
};

// relu(sc * x + sh) on 4 consecutive channels
__device__ __forceinline__ void xg_bn4(float* r, const float4& sc, const float4& sh) {
  r[0] = fmaxf(fmaf(sc.x, r[0], sh.x), 0.0f); r[1] = fmaxf(fmaf(sc.y, r[1], sh.y), 0.0f);
  r[2] = fmaxf(fmaf(sc.z, r[2], sh.z), 0.0f); r[3] = fmaxf(fmaf(sc.w, r[3], sh.w), 0.0f);
}

// row of the tile handled by (warp w of the operand's 4, iteration i, lane) in the UNIT_K mapping: a warp-wide 128-bit
// load covers 4 rows x 128 B (coalesced), and the 4 rows are r, r+4 (lanes 0-15) and r+1, r+5 (lanes 16-31) of an 8-row
// swizzle atom: the 64-bit shared stores are issued per half-warp, and rows r / r+4 land in opposite 64-byte halves of
// the 128-byte bank span (ncu: 37-52 % of the store wavefronts were bank conflicts with rows r, r+1 in one half-warp).
__device__ __forceinline__ int xg_row_unit_k(int w, int i, int lane) {
  const int rsub = lane >> 3;
  return w * 32 + (i >> 1) * 8 + (i & 1) * 2 + (rsub & 1) * 4 + (rsub >> 1);
}

// One stage (128 rows x 32 k) of an operand tile into registers.  `p` is this thread's pointer for the stage:
//   UNIT_K  : &P[(row0 + xg_row_unit_k(w, 0, lane)) * ld + k0 + 4 * (lane & 7)]; the 8 rows of the thread are
//             +0, +2, +8, +10, +16, +18, +24, +26 rows from there (128-bit loads, a warp covers 4 rows x 128 B)
//   !UNIT_K : &P[k0 * ld + row0 + w * 32 + lane]; 32 consecutive k (a warp reads 32 consecutive floats of one k)
// `full` = every row and k of the stage is in range and (UNIT_K) 128-bit loads are allowed: no predicates.
template <bool UNIT_K>
__device__ __forceinline__ void xg_load(const float* __restrict__ p, long long ld, bool full, int row, int nrows, int k, int kend,
                                        float (&r)[32]) {
  if (UNIT_K) {
    if (full) {
      const long long ld2 = 2 * ld, ld6 = 6 * ld;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p));
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        p += (i & 1) ? ld6 : ld2;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int ro = (i >> 1) * 8 + (i & 1) * 2;
        const float* q = p + (long long)ro * ld;
#pragma unroll
        for (int e = 0; e < 4; ++e) r[4 * i + e] = (row + ro < nrows && k + e < kend) ? __ldg(q + e) : 0.0f;
      }
    }
  } else {
    if (full) {
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) { r[kk] = __ldg(p); p += ld; }
    } else {
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) r[kk] = (row < nrows && k + kk < kend) ? __ldg(p + (long long)kk * ld) : 0.0f;
    }
  }
}

// x = h + m + l exactly, each piece a bf16 held in the top half of a 32-bit word
__device__ __forceinline__ void xg_split(float x, uint32_t& h, uint32_t& m, uint32_t& l) {
  h = __float_as_uint(x) & 0xffff0000u;
  const float r = x - __uint_as_float(h);
  m = __float_as_uint(r) & 0xffff0000u;
  l = __float_as_uint(r - __uint_as_float(m));
}
__device__ __forceinline__ uint32_t xg_pack(uint32_t lo, uint32_t hi) { return __byte_perm(lo, hi, 0x7632); }   // top halves

__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

// registers of xg_load -> the three bf16 images of stage s (img = shared address of image 1; 2 and 3 follow)
__device__ __forceinline__ uint32_t xg_pack_rn(float lo, float hi) {      // round-to-nearest bf16x2 (lo in the low half)
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// PARTS images of 4 consecutive fp32 values as packed bf16 pairs: w[p][0] = (x0, x1), w[p][1] = (x2, x3) of image p.
//   PARTS 1: round to nearest (one product per MAC, the approximate `bf16` engine)
//   PARTS 2: x = h + m + O(2^-18 x), both pieces rounded to nearest (three products per MAC, the `tc2` engine)
//   PARTS 3: x = h + m + l exactly, by truncation (six products per MAC, the fp32-accurate `tc` engine)
template <int PARTS>
__device__ __forceinline__ void xg_split4(const float* x, uint32_t (&w)[3][2]) {
  if (PARTS == 1) {
    w[0][0] = xg_pack_rn(x[0], x[1]); w[0][1] = xg_pack_rn(x[2], x[3]);
  } else if (PARTS == 2) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const uint32_t h = xg_pack_rn(x[2 * q], x[2 * q + 1]);
      w[0][q] = h;
      w[1][q] = xg_pack_rn(x[2 * q] - __uint_as_float(h << 16), x[2 * q + 1] - __uint_as_float(h & 0xffff0000u));
    }
  } else {
    uint32_t h[4], m[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) xg_split(x[e], h[e], m[e], l[e]);
    w[0][0] = xg_pack(h[0], h[1]); w[0][1] = xg_pack(h[2], h[3]);
    w[1][0] = xg_pack(m[0], m[1]); w[1][1] = xg_pack(m[2], m[3]);
    w[2][0] = xg_pack(l[0], l[1]); w[2][1] = xg_pack(l[2], l[3]);
  }
}

template <bool UNIT_K, int PARTS>
__device__ __forceinline__ void xg_split_store(uint32_t img, int s, int w, int lane, const float (&r)[32]) {
  if (UNIT_K) {
    const int c = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t addr = img + sw128_offset((uint32_t)xg_row_unit_k(w, i, lane), (uint32_t)(4 * s + (c >> 1))) + (uint32_t)(c & 1) * 8u;
      uint32_t w[3][2];
      xg_split4<PARTS>(&r[4 * i], w);
#pragma unroll
      for (int p = 0; p < PARTS; ++p) st_shared_v2(addr + (uint32_t)p * kXgImage, w[p][0], w[p][1]);
    }
  } else {
    const uint32_t row = (uint32_t)(w * 32 + lane);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t addr = img + sw128_offset(row, (uint32_t)(4 * s + j));
      uint32_t w0[3][2], w1[3][2];
      xg_split4<PARTS>(&r[8 * j], w0);
      xg_split4<PARTS>(&r[8 * j + 4], w1);
#pragma unroll
      for (int p = 0; p < PARTS; ++p) st_shared_v4(addr + (uint32_t)p * kXgImage, w0[p][0], w0[p][1], w1[p][0], w1[p][1]);
    }
  }
}

struct XgTracer {          // (clock64 << 8 | tag) per role, recorded by one thread of the middle CTA of the grid
  unsigned long long* buf; int n;
  __device__ __forceinline__ void init(unsigned long long* base, int role, bool who) {
    buf = (base != nullptr && who && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0) ? base + (size_t)role * kTraceSlots : nullptr; n = 0;
  }
  __device__ __forceinline__ void mark(int tag) {
    if (buf != nullptr && n < kTraceSlots) buf[n++] = ((unsigned long long)clock64() << 8) | (unsigned long long)(tag & 0xff);
  }
};

__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {     // many threads poll: yield issue slots
  while (!mbar_try_wait(bar, parity)) __nanosleep(32);
}

// One operand's loader warp: stage it+1 is in flight in registers while stage it is split and stored (two register
// sets, ping-pong).  full0 / empty0 = the stage-0 barriers (stage 1 follows at +8 bytes).
template <bool UNIT_K, int PARTS>
__device__ __forceinline__ void xg_loader(const float* __restrict__ P, long long ld, int row0, int nrows, int kbeg, int kend, bool vec,
                                          uint32_t img, int w, int lane, int nst, uint32_t full0, uint32_t empty0, XgTracer& tr,
                                          const float* __restrict__ tsc = nullptr, const float* __restrict__ tsh = nullptr,
                                          bool remote_full = false) {      // remote_full: the full barriers live in cluster rank 0 (CTA-pair kernel)
  const int row = row0 + (UNIT_K ? xg_row_unit_k(w, 0, lane) : w * 32 + lane);      // first (or only) row of this thread
  const int kofs = UNIT_K ? 4 * (lane & 7) : 0;
  const float* p = UNIT_K ? P + (long long)row * ld + kbeg + kofs : P + (long long)kbeg * ld + row;
  const long long pstep = UNIT_K ? (long long)kXgBK : (long long)kXgBK * ld;
  const bool rows_full = (row0 + w * 32 + 32 <= nrows) && (UNIT_K ? vec : true);
  auto load = [&](int it, float (&r)[32]) {
    const int k0 = kbeg + it * kXgBK;
    xg_load<UNIT_K>(p + (long long)it * pstep, ld, rows_full && (k0 + kXgBK <= kend), row, nrows, k0 + kofs, kend, r);
  };
  // lazy BN of a row-contiguous operand (wgrad: A = X^T, row = channel of X): one scale / shift pair per thread; k past the
  // end of the range and rows past the operand stay zero
  const bool bn = !UNIT_K && tsc != nullptr;
  const float bsc = (bn && row < nrows) ? __ldg(tsc + row) : 0.0f, bsh = (bn && row < nrows) ? __ldg(tsh + row) : 0.0f;
  auto emit = [&](int it, float (&r)[32]) {
    const int s = it & 1;
    if (bn) {
      const int k0 = kbeg + it * kXgBK;
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) r[kk] = (row < nrows && k0 + kk < kend) ? fmaxf(fmaf(bsc, r[kk], bsh), 0.0f) : 0.0f;
    }
    if (UNIT_K && tsc != nullptr) {      // lazy BN of a k-contiguous operand (forward): this thread's 4 channels of the stage (K % 32 == 0)
      const int k = kbeg + it * kXgBK + kofs;
      const float4 sc = make_float4(__ldg(tsc + k), __ldg(tsc + k + 1), __ldg(tsc + k + 2), __ldg(tsc + k + 3));
      const float4 sh = make_float4(__ldg(tsh + k), __ldg(tsh + k + 1), __ldg(tsh + k + 2), __ldg(tsh + k + 3));
#pragma unroll
      for (int i = 0; i < 8; ++i) xg_bn4(&r[4 * i], sc, sh);
    }
    mbar_wait_backoff(empty0 + 8u * s, ((it >> 1) & 1) ^ 1);
    tr.mark(0x20);
    xg_split_store<UNIT_K, PARTS>(img, s, w, lane, r);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) { if (remote_full) mbar_arrive_remote(full0 + 8u * s, 0); else mbar_arrive(full0 + 8u * s); }
    tr.mark(0x30);
  };
  // kXgSets register sets: stage `it` lives in set it % kXgSets; kXgSets - 1 stages are in flight while one is consumed
  float r[kXgSets][32];
#pragma unroll
  for (int j = 0; j < kXgSets - 1; ++j) if (j < nst) load(j, r[j]);
  tr.mark(0x10);
  for (int it0 = 0; it0 < nst; it0 += kXgSets) {
#pragma unroll
    for (int j = 0; j < kXgSets; ++j) {
      const int it = it0 + j;
      if (it < nst) {
        if (it + kXgSets - 1 < nst) load(it + kXgSets - 1, r[(j + kXgSets - 1) % kXgSets]);
        tr.mark(0x10);
        emit(it, r[j]);
      }
    }
  }
}

// ---- pre-split B (forward / dgrad: B is the small weight matrix, the same for every row tile) -------------------------
// A pre-pass splits B once per call into the exact shared-memory images ([128 rows x 64 k] SW128 blocks, 3 parts) in a
// caller-provided workspace; the main kernel then brings its B stage in with 16-byte cp.async copies (no registers, no
// split arithmetic, no per-CTA re-splitting of the same weights) and all 8 loader warps split A, 16 values per thread
// and stage instead of 32.
constexpr size_t kXgPreBlockBytes = 3 * (size_t)kXgImage;       // one (n tile, k block): 3 parts x 16 KB
inline size_t xg_pre_bytes(int N, int K) { return (size_t)((N + kXgBN - 1) / kXgBN) * ((K + 63) / 64) * kXgPreBlockBytes; }

// grid (k blocks, n tiles, 4): block z handles k groups z*2, z*2+1 (8 k each) of the 64-wide block; a thread owns one
// row and 8 consecutive k, i.e. one 16-byte chunk of each image.  (The first version wrote 2 bytes per store from 64 blocks
// and took as long as the FC-head GEMMs it served: 29 us per call in the cfg3 launch list.)
__global__ void __launch_bounds__(256) xg_presplit_kernel(const float* __restrict__ B, long long ldb, int unit_k, int N, int K,
                                                         int parts, uint8_t* __restrict__ ws) {
  const int kb = blockIdx.x, nt = blockIdx.y;
  uint8_t* dst = ws + ((size_t)nt * gridDim.x + kb) * kXgPreBlockBytes;
  const int r = threadIdx.x & 127, kg = blockIdx.z * 2 + (threadIdx.x >> 7);       // row, 8-k group (= 16-byte chunk)
  const int n = nt * kXgBN + r, k0 = kb * 64 + kg * 8;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = k0 + e;
    v[e] = (n < N && k < K) ? (unit_k ? B[(long long)n * ldb + k] : B[(long long)k * ldb + n]) : 0.0f;
  }
  const uint32_t off = sw128_offset((uint32_t)r, (uint32_t)kg);
  uint32_t w0[3][2], w1[3][2];
  if (parts == 1) { xg_split4<1>(v, w0); xg_split4<1>(v + 4, w1); }
  else if (parts == 2) { xg_split4<2>(v, w0); xg_split4<2>(v + 4, w1); }
  else { xg_split4<3>(v, w0); xg_split4<3>(v + 4, w1); }
  for (int p = 0; p < parts; ++p)
    *reinterpret_cast<uint4*>(dst + (size_t)p * kXgImage + off) = make_uint4(w0[p][0], w0[p][1], w1[p][0], w1[p][1]);
}

__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {      // contiguous range -> L2 (16-byte granularity)
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// loader warp w of 8 (k-contiguous A, pre-split B): rows w*16 .. w*16+15 of the A tile + 1/8 of the B stage copy
template <int PARTS>
__device__ __forceinline__ void xg_loader_pre(const float* __restrict__ P, long long ld, int row0, int nrows, int kend, bool vec,
                                              uint32_t a_img, uint32_t b_img, const uint8_t* __restrict__ bsrc, int w, int lane, int nst,
                                              uint32_t full0, uint32_t empty0, XgTracer& tr,
                                              uint32_t tsc = 0, uint32_t tsh = 0) {      // shared-memory tables of the lazy BN map (0: none)
  const int rsub = lane >> 3, c = lane & 7;
  auto rowof = [&](int i) { return w * 16 + (i >> 1) * 8 + (i & 1) * 2 + (rsub & 1) * 4 + (rsub >> 1); };
  const int row = row0 + rowof(0), kofs = 4 * c;
  const float* p = P + (long long)row * ld + kofs;
  const bool rows_full = (row0 + w * 16 + 16 <= nrows) && vec;
  const long long ld2 = 2 * ld, ld6 = 6 * ld;
  auto load = [&](int it, float (&r)[16]) {
    const int k0 = it * kXgBK;
    const float* q = p + k0;
    if (XG_DBG & 1) {
#pragma unroll
      for (int e = 0; e < 16; ++e) r[e] = (float)(it + e);
    } else if (rows_full && k0 + kXgBK <= kend) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(q));
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
        q += (i & 1) ? ld6 : ld2;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ro = (i >> 1) * 8 + (i & 1) * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) r[4 * i + e] = (row + ro < nrows && k0 + kofs + e < kend) ? __ldg(q + (long long)ro * ld + e) : 0.0f;
      }
    }
  };
  const int tid = w * 32 + lane;
  auto emit = [&](int it, float (&r)[16]) {
    const int s = it & 1;
    if (tsc != 0) {            // lazy BN: this thread's 4 channels of the stage
      const float4 sc = ld_shared_f4(tsc + 4u * (uint32_t)(it * kXgBK + kofs));
      const float4 sh = ld_shared_f4(tsh + 4u * (uint32_t)(it * kXgBK + kofs));
#pragma unroll
      for (int i = 0; i < 4; ++i) xg_bn4(&r[4 * i], sc, sh);
    }
    mbar_wait_backoff(empty0 + 8u * s, ((it >> 1) & 1) ^ 1);
    tr.mark(0x20);
    {   // B stage: PARTS x 128 rows x 4 chunks of 16 B, same swizzled offsets in the workspace block and in shared memory
      const uint8_t* blk = bsrc + (size_t)(it >> 1) * kXgPreBlockBytes;
#pragma unroll
      for (int j = 0; j < ((XG_DBG & 8) ? 0 : PARTS * 2); ++j) {
        const int id = tid + 256 * j, part = id >> 9, rem = id & 511;
        const uint32_t off = (uint32_t)part * kXgImage + sw128_offset((uint32_t)(rem >> 2), (uint32_t)(4 * s + (rem & 3)));
        cp_async16(b_img + off, blk + off);
      }
    }
#pragma unroll
    for (int i = 0; i < ((XG_DBG & 2) ? 0 : 4); ++i) {
      const uint32_t addr = a_img + sw128_offset((uint32_t)rowof(i), (uint32_t)(4 * s + (c >> 1))) + (uint32_t)(c & 1) * 8u;
      uint32_t w[3][2];
      xg_split4<PARTS>(&r[4 * i], w);
#pragma unroll
      for (int p = 0; p < PARTS; ++p) st_shared_v2(addr + (uint32_t)p * kXgImage, w[p][0], w[p][1]);
    }
    if (XG_DBG & 2) {      // keep the loads alive without the stores
      uint32_t x = 0;
#pragma unroll
      for (int e = 0; e < 16; ++e) x ^= __float_as_uint(r[e]);
      if (x == 0xdeadbeefu) st_shared_v2(a_img, x, x);
    }
    cp_async_wait_all();
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(full0 + 8u * s);
    tr.mark(0x30);
  };
  float r[kXgPreSets][16];
#pragma unroll
  for (int j = 0; j < kXgPreSets - 1; ++j) if (j < nst) load(j, r[j]);
  for (int it0 = 0; it0 < nst; it0 += kXgPreSets) {
#pragma unroll
    for (int j = 0; j < kXgPreSets; ++j) {
      const int it = it0 + j;
      if (it < nst) {
        if (it + kXgPreSets - 1 < nst) load(it + kXgPreSets - 1, r[(j + kXgPreSets - 1) % kXgPreSets]);
        emit(it, r[j]);
      }
    }
  }
}

// the partial products of one K = 16 step (warp-convergent): small accumulator first, the leading product last
template <int PARTS>
__device__ __forceinline__ void xg_issue(uint32_t d_main, uint32_t d_small, uint64_t a1, uint64_t a2, uint64_t a3, uint64_t b1, uint64_t b2,
                                         uint64_t b3, uint32_t idesc, uint32_t accum) {
  if (PARTS == 3) {
    umma_bf16_w(d_small, a3, b1, idesc, accum);
    umma_bf16_w(d_small, a1, b3, idesc, 1);
    umma_bf16_w(d_small, a2, b2, idesc, 1);
    umma_bf16_w(d_small, a2, b1, idesc, 1);
    umma_bf16_w(d_small, a1, b2, idesc, 1);
  } else if (PARTS == 2) {
    umma_bf16_w(d_small, a2, b1, idesc, accum);
    umma_bf16_w(d_small, a1, b2, idesc, 1);
  }
  umma_bf16_w(d_main, a1, b1, idesc, accum);
}

struct XgTile {
  uint32_t sbase, tmem_base;
  int m0, n0, warp, lane;
};

// Runs the main loop of one tile.  Returns in the 8 loader warps once both accumulators are complete in TMEM
// (main: columns [0,128), small: [128,256) of the allocation; lane = tile row); warp 8 returns immediately after its
// last commit.  Every thread must then call xg_finish().
template <bool A_UNIT_K, bool B_UNIT_K, int PARTS, bool B_PRE = false>
__device__ __forceinline__ void xg_mainloop(const XgOperands& o, uint8_t* smem_raw, XgTile& t) {
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = sbase + kXgBars;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (2 + s); };
  const uint32_t acc_full = bar0 + 32u;
  const uint32_t tmem_slot = bar0 + 40u;
  const int m0 = (int)(blockIdx.x / o.ntn) * kXgBM, n0 = (int)(blockIdx.x % o.ntn) * kXgBN;
  const int kbeg = blockIdx.y * o.kchunk, kend = min(o.K, kbeg + o.kchunk);
  const int nst = (kend - kbeg + kXgBK - 1) / kXgBK;

  if (threadIdx.x == 0) {
    mbar_init(full(0), 8); mbar_init(full(1), 8);
    mbar_init(empty(0), 1); mbar_init(empty(1), 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc<256>(tmem_slot);
  if (B_PRE && o.a_scale != nullptr) {      // lazy BN tables of the A operand (K <= kXgLazyMaxK, checked by the host)
    float* tab = reinterpret_cast<float*>(smem + kXgParams);
    for (int i = threadIdx.x; i < o.K; i += kXgThreads) { tab[i] = __ldg(o.a_scale + i); tab[o.K + i] = __ldg(o.a_shift + i); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + kXgBars + 40);
  t.sbase = sbase; t.tmem_base = tmem_base; t.m0 = m0; t.n0 = n0; t.warp = warp; t.lane = lane;
  XgTracer tr;
  tr.init(o.trace, warp == 0 ? 0 : (warp == 4 ? 1 : 2), lane == 0 && (warp == 0 || warp == 4 || warp == 8));
  tr.mark(0x02);

  if (warp < 8) {
    // ---------------------------------------------------------------- loaders: global fp32 -> registers -> 3 bf16 images
    if (B_PRE) xg_loader_pre<PARTS>(o.A, o.lda, m0, o.M, kend, o.vecA != 0, sbase, sbase + 3u * kXgImage,
                                    o.bpre + (size_t)(blockIdx.x % o.ntn) * o.nkb * kXgPreBlockBytes, warp, lane, nst, full(0), empty(0), tr,
                                    o.a_scale != nullptr ? sbase + kXgParams : 0u, sbase + kXgParams + 4u * (uint32_t)o.K);
    else if (warp < 4) xg_loader<A_UNIT_K, PARTS>(o.A, o.lda, m0, o.M, kbeg, kend, o.vecA != 0, sbase, warp & 3, lane, nst, full(0), empty(0), tr,
                                                  o.a_scale, o.a_shift);
    else xg_loader<B_UNIT_K, PARTS>(o.B, o.ldb, n0, o.N, kbeg, kend, o.vecB != 0, sbase + 3u * kXgImage, warp & 3, lane, nst, full(0), empty(0), tr);
    mbar_wait_backoff(acc_full, 0);
    tc_fence_after();
    tr.mark(0x40);
  } else {
    // ---------------------------------------------------------------- MMA issue (warp-convergent, one elected lane)
    const int ncols = min(kXgBN, ((o.N - n0) + 15) & ~15);
    const uint32_t idesc = make_idesc_bf16(128, ncols);
    const uint32_t d_main = tmem_base, d_small = tmem_base + 128;
    const uint32_t a_img = sbase, b_img = sbase + 3 * kXgImage;
    for (int it = 0; it < nst; ++it) {
      const int s = it & 1;
      mbar_wait_w(full(s), (it >> 1) & 1);
      tc_fence_after();
      tr.mark(0x10);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t off = (uint32_t)(2 * s + ks) * 32u;
        const uint64_t a1 = make_sdesc_k128(a_img + off), a2 = make_sdesc_k128(a_img + kXgImage + off),
                       a3 = make_sdesc_k128(a_img + 2 * kXgImage + off);
        const uint64_t b1 = make_sdesc_k128(b_img + off), b2 = make_sdesc_k128(b_img + kXgImage + off),
                       b3 = make_sdesc_k128(b_img + 2 * kXgImage + off);
        const uint32_t first = (it | ks) != 0;
        if (XG_DBG & 4) continue;
        xg_issue<PARTS>(d_main, d_small, a1, a2, a3, b1, b2, b3, idesc, first);
      }
      umma_commit_w(empty(s));
      tr.mark(0x20);
    }
    umma_commit_w(acc_full);
  }
}

// 32 accumulator columns [c0, c0+32) of this thread's row: main + small
template <int PARTS>
__device__ __forceinline__ void xg_acc32(const XgTile& t, int c0, float (&v)[32]) {
  const uint32_t taddr = t.tmem_base + ((uint32_t)((t.warp & 3) * 32) << 16) + (uint32_t)c0;
  uint32_t va[32];
  tmem_ld32(taddr, va);
  if (PARTS >= 2) {
    uint32_t vb[32];
    tmem_ld32(taddr + 128, vb);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(va[j]) + __uint_as_float(vb[j]);
  } else {
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(va[j]);
  }
}

__device__ __forceinline__ void xg_finish(const XgTile& t) {
  tc_fence_before();
  __syncthreads();
  if (t.warp == 8) tmem_dealloc<256>(t.tmem_base);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// The warp's 32 x 32 chunk (thread = row, v = 32 consecutive columns) goes through a warp-private shared-memory patch
// (the operand images are dead once acc_full has fired) and leaves as full 128-byte row segments: 8 lanes per row, 4 rows
// per store instruction.  Storing straight from the accumulator layout (a 16-byte piece of 32 different lines per
// instruction) cost 20 - 47 % of the kernel on the output-heavy forward layers (measured by removing the epilogue).
// Patch rows are 144 bytes apart, so both the 128-bit writes (thread = row) and reads (8 lanes = one row) are
// conflict-free.
constexpr uint32_t kXgPatchRow = 144, kXgPatchBytes = 32 * kXgPatchRow;
// MODE 1: the column sums of (x - shift[col]) and of its square over the chunk's valid rows are added to st_sum / st_sq
// (training-mode batch norm: the statistics pass over the layer output is fused here; shift = row 0 of the output keeps
// E[d^2] - E[d]^2 well conditioned).  Each lane sums its 8 rows of 4 columns with packed fp32 arithmetic (FFMA2: the
// epilogue warps are the limit of the K <= 128 kernels, and the scalar version of this block cost them 30 %), two shuffles
// fold the 4 row groups.  MODE 2: additionally the column max / min of the chunk's valid rows as ordered keys, merged into
// pool_max / pool_min + pool_off (the group's row of the [groups, N] key arrays); C == null skips the store.
template <bool ATOMIC, int MODE = 0>
__device__ __forceinline__ void xg_store_chunk(const XgTile& t, float* __restrict__ C, long long ldc, int M, int N, int row0, int col0,
                                               const float (&v)[32], float* __restrict__ st_sum = nullptr,
                                               float* __restrict__ st_sq = nullptr, const float* __restrict__ st_shift = nullptr,
                                               unsigned* __restrict__ pool_max = nullptr, unsigned* __restrict__ pool_min = nullptr,
                                               size_t pool_off = 0, float4 sfv = make_float4(0.f, 0.f, 0.f, 0.f)) {
  const uint32_t patch = t.sbase + (uint32_t)t.warp * kXgPatchBytes;
  __syncwarp();                                          // the previous chunk's reads are done
#pragma unroll
  for (int j = 0; j < 8; ++j)
    st_shared_v4(patch + (uint32_t)t.lane * kXgPatchRow + 16u * j, __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                 __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
  __syncwarp();
  const int cq = (t.lane & 7) * 4, gn = col0 + cq;
  // packed accumulators: columns (0, 1) and (2, 3) of this lane's quad; the shift enters negated (sfv fetched by the caller)
  unsigned long long ss01 = 0ull, ss23 = 0ull, sq01 = 0ull, sq23 = 0ull;
  const unsigned long long nsf01 = xg_pk2(-sfv.x, -sfv.y), nsf23 = xg_pk2(-sfv.z, -sfv.w);
  float pmx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, pmn[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = (t.lane >> 3) + 4 * i, gm = row0 + r;
    const float4 x = ld_shared_f4(patch + (uint32_t)r * kXgPatchRow + 4u * cq);
    if (gm >= M || gn >= N) continue;
    if (MODE >= 1) {
      const unsigned long long d01 = xg_add2(xg_pk2(x.x, x.y), nsf01), d23 = xg_add2(xg_pk2(x.z, x.w), nsf23);
      ss01 = xg_add2(ss01, d01); ss23 = xg_add2(ss23, d23);
      sq01 = xg_fma2(d01, d01, sq01); sq23 = xg_fma2(d23, d23, sq23);
    }
    if (MODE == 2) {
      pmx[0] = fmaxf(pmx[0], x.x); pmx[1] = fmaxf(pmx[1], x.y); pmx[2] = fmaxf(pmx[2], x.z); pmx[3] = fmaxf(pmx[3], x.w);
      pmn[0] = fminf(pmn[0], x.x); pmn[1] = fminf(pmn[1], x.y); pmn[2] = fminf(pmn[2], x.z); pmn[3] = fminf(pmn[3], x.w);
      if (C == nullptr) continue;
    }
    float* c = C + (long long)gm * ldc + gn;
    if (gn + 4 <= N && (((uintptr_t)c) & 15) == 0) {
      if (ATOMIC) red_add_v4(c, x.x, x.y, x.z, x.w);
      else *reinterpret_cast<float4*>(c) = x;
    } else {
      const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (gn + e < N) {
          if (ATOMIC) atomicAdd(c + e, xs[e]);
          else c[e] = xs[e];
        }
      }
    }
  }
  if (MODE >= 1) {
    float ss[4], sq[4];
    xg_upk2(ss01, ss[0], ss[1]); xg_upk2(ss23, ss[2], ss[3]); xg_upk2(sq01, sq[0], sq[1]); xg_upk2(sq23, sq[2], sq[3]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ss[e] += __shfl_xor_sync(0xffffffffu, ss[e], 8); ss[e] += __shfl_xor_sync(0xffffffffu, ss[e], 16);
      sq[e] += __shfl_xor_sync(0xffffffffu, sq[e], 8); sq[e] += __shfl_xor_sync(0xffffffffu, sq[e], 16);
    }
    if (MODE == 2) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        pmx[e] = fmaxf(pmx[e], __shfl_xor_sync(0xffffffffu, pmx[e], 8)); pmx[e] = fmaxf(pmx[e], __shfl_xor_sync(0xffffffffu, pmx[e], 16));
        pmn[e] = fminf(pmn[e], __shfl_xor_sync(0xffffffffu, pmn[e], 8)); pmn[e] = fminf(pmn[e], __shfl_xor_sync(0xffffffffu, pmn[e], 16));
      }
      if (t.lane < 8 && pmx[0] >= pmn[0]) {         // at least one valid row in the chunk
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (gn + e < N) {
            atomicMax(pool_max + pool_off + gn + e, f32_ordered(pmx[e]));
            atomicMin(pool_min + pool_off + gn + e, f32_ordered(pmn[e]));
          }
      }
    }
    if (t.lane < 8 && gn < N) {
      if (gn + 4 <= N && ((((uintptr_t)(st_sum + gn)) | ((uintptr_t)(st_sq + gn))) & 15) == 0) {
        red_add_v4(st_sum + gn, ss[0], ss[1], ss[2], ss[3]);
        red_add_v4(st_sq + gn, sq[0], sq[1], sq[2], sq[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (gn + e < N) { atomicAdd(st_sum + gn + e, ss[e]); atomicAdd(st_sq + gn + e, sq[e]); }
      }
    }
  }
}

#ifdef T3D_XGEMM_WITH_EPILOGUES
// ---- t3d_gemm_f32 (same contract as gemm_f32_kernel, train_ops.cuh): C = A.B (+ bias by split 0); split-K partial tiles
// are added into a zero-initialised C with vector reductions.
// t.warp: 0-7 in the one-tile kernels (lane quarter = warp & 3, column half = warp >> 2, 2 chunks), 0-3 in the persistent
// kernel (4 chunks: the whole row)
template <int PARTS>
__device__ __forceinline__ void xg_epilogue_gemm(const GemmArgs& a, const XgTile& t, int nchunks = 2, bool first_split = true) {
  if (t.warp < 8 && !(XG_DBG & 16)) {
    const int row0 = t.m0 + (t.warp & 3) * 32;
#pragma unroll 1
    for (int ch = 0; ch < nchunks; ++ch) {
      const int c0 = (t.warp >> 2) * (32 * nchunks) + ch * 32;      // warps 4-7 of the one-tile kernels: the second column half
      float4 sfv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.st_sum != nullptr && a.st_shift != nullptr) {     // this lane's 4 statistic shifts, in flight during the TMEM load
        const int gs = t.n0 + c0 + (t.lane & 7) * 4;
        sfv.x = gs < a.N ? __ldg(a.st_shift + gs) : 0.f; sfv.y = gs + 1 < a.N ? __ldg(a.st_shift + gs + 1) : 0.f;
        sfv.z = gs + 2 < a.N ? __ldg(a.st_shift + gs + 2) : 0.f; sfv.w = gs + 3 < a.N ? __ldg(a.st_shift + gs + 3) : 0.f;
      }
      float v[32];
      xg_acc32<PARTS>(t, c0, v);
      if (a.bias && first_split) {
        float q[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) q[j] = (t.n0 + c0 + j < a.N) ? __ldg(a.bias + t.n0 + c0 + j) : 0.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += q[j];
      }
      if (a.splitk > 1) xg_store_chunk<true>(t, a.C, a.ldc, a.M, a.N, row0, t.n0 + c0, v);
      else if (a.pool_max != nullptr)
        xg_store_chunk<false, 2>(t, a.C, a.ldc, a.M, a.N, row0, t.n0 + c0, v, a.st_sum, a.st_sq, a.st_shift, a.pool_max, a.pool_min,
                                 (size_t)(t.m0 / a.pool_rows) * a.N, sfv);
      else if (a.st_sum != nullptr)
        xg_store_chunk<false, 1>(t, a.C, a.ldc, a.M, a.N, row0, t.n0 + c0, v, a.st_sum, a.st_sq, a.st_shift, nullptr, nullptr, 0, sfv);
      else xg_store_chunk<false>(t, a.C, a.ldc, a.M, a.N, row0, t.n0 + c0, v);
    }
  }
}

template <bool A_UNIT_K, bool B_UNIT_K, int PARTS>
__global__ void __launch_bounds__(kXgThreads, kXgMinBlocks) xgemm_kernel(const GemmArgs a, const XgOperands o) {
  extern __shared__ uint8_t xg_smem[];
  XgTile t;
  xg_mainloop<A_UNIT_K, B_UNIT_K, PARTS>(o, xg_smem, t);
  xg_epilogue_gemm<PARTS>(a, t, 2, blockIdx.y == 0);
  xg_finish(t);
}
// k-contiguous A, pre-split B (forward: B = W; dgrad: B = W^T)
template <int PARTS>
__global__ void __launch_bounds__(kXgThreads, kXgMinBlocks) xgemm_pre_kernel(const GemmArgs a, const XgOperands o) {
  extern __shared__ uint8_t xg_smem[];
  XgTile t;
  xg_mainloop<true, false, PARTS, true>(o, xg_smem, t);
  xg_epilogue_gemm<PARTS>(a, t, 2, blockIdx.y == 0);
  xg_finish(t);
}

// ---- CTA-pair variant: one 256 x 256 output tile per cluster of two CTAs (tcgen05 cta_group::2) -----------------------------
// For the two-piece and one-piece engines (PARTS <= 2).  CTA r of the pair loads, splits and keeps in ITS shared memory the
// A rows [m0 + 128 r, +128) and the B columns [n0 + 128 r, +128) of every stage; the leader's MMA warp issues M = 256 x N = 256
// instructions that read both shared memories and fill a [128 x 256] fp32 accumulator in each CTA's TMEM.  Against the
// 128 x 128 tile of the kernels above every operand value loaded and split feeds twice the tensor-core work (these kernels
// are bound by their loaders and by the L2 -> SM operand traffic, not by the tensor pipe: DESIGN.md section 4b), and B needs
// no pre-split pass.  One accumulator (the products of the low pieces are summed into the main one): 256 TMEM columns per
// CTA, so two clusters share an SM pair and one's epilogue runs under the other's main loop.  The loaders of the peer arrive
// remotely on the leader's full barriers; the leader's commits are multicast to both CTAs.
// main loop of one pair tile; returns in every warp once both CTAs' accumulators are complete (acc_full).  Every thread then
// runs its epilogue and calls xg_pair_finish().
template <bool A_UNIT_K, bool B_UNIT_K, int PARTS>
__device__ __forceinline__ void xg_pair_mainloop(const XgOperands& o, uint8_t* smem_raw, XgTile& t, XgTracer& tr) {
  static_assert(PARTS <= 2, "xgemm_pair: one accumulator");
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t bar0 = sbase + kXgBars;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (2 + s); };
  const uint32_t acc_full = bar0 + 32u;
  const uint32_t tmem_slot = bar0 + 40u;
  const int ptile = (int)(blockIdx.x >> 1);                      // o.ntn = number of 256-wide column tiles
  const int m0 = (ptile / o.ntn) * 256 + 128 * (int)rank, n0 = (ptile % o.ntn) * 256;
  const int kbeg = blockIdx.y * o.kchunk, kend = min(o.K, kbeg + o.kchunk);
  const int nst = (kend - kbeg + kXgBK - 1) / kXgBK;

  if (threadIdx.x == 0) {
    mbar_init(full(0), 16); mbar_init(full(1), 16);             // 8 loader warps of each CTA (used in the leader only)
    mbar_init(empty(0), 1); mbar_init(empty(1), 1);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  const long long t_entry = clock64();
  if (warp == 8) tmem_alloc_pair<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + kXgBars + 40);
  t.sbase = sbase; t.tmem_base = tmem_base; t.m0 = m0; t.n0 = n0; t.warp = warp; t.lane = lane;
  tr.init(o.trace, warp == 0 ? 0 : (warp == 4 ? 1 : 2), lane == 0 && (warp == 0 || warp == 4 || warp == 8));
  if (tr.buf != nullptr) tr.buf[tr.n++] = ((unsigned long long)t_entry << 8) | 0x01;      // kernel entry
  tr.mark(0x02);

  if (warp < 8) {
    if (warp < 4) xg_loader<A_UNIT_K, PARTS>(o.A, o.lda, m0, o.M, kbeg, kend, o.vecA != 0, sbase, warp, lane, nst, full(0), empty(0), tr,
                                             o.a_scale, o.a_shift, rank != 0);
    else xg_loader<B_UNIT_K, PARTS>(o.B, o.ldb, n0 + 128 * (int)rank, o.N, kbeg, kend, o.vecB != 0, sbase + 3u * kXgImage, warp & 3, lane, nst,
                                    full(0), empty(0), tr, nullptr, nullptr, rank != 0);
    mbar_wait_backoff(acc_full, 0);
    tc_fence_after();
    tr.mark(0x40);
  } else if (rank == 0) {
    // ---------------------------------------------------------------- MMA issue (leader; warp-convergent, one elected lane)
    const uint32_t idesc = make_idesc_bf16(256, 256);
    const uint32_t a_img = sbase, b_img = sbase + 3 * kXgImage;
    for (int it = 0; it < nst; ++it) {
      const int s = it & 1;
      mbar_wait_w(full(s), (it >> 1) & 1);
      tc_fence_after();
      tr.mark(0x10);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t off = (uint32_t)(2 * s + ks) * 32u;
        const uint64_t a1 = make_sdesc_k128(a_img + off), a2 = make_sdesc_k128(a_img + kXgImage + off);
        const uint64_t b1 = make_sdesc_k128(b_img + off), b2 = make_sdesc_k128(b_img + kXgImage + off);
        const uint32_t accum = (it | ks) != 0;
        if (PARTS == 2) {
          umma_bf16_pair_w(tmem_base, a2, b1, idesc, accum);
          umma_bf16_pair_w(tmem_base, a1, b2, idesc, 1);
          umma_bf16_pair_w(tmem_base, a1, b1, idesc, 1);
        } else {
          umma_bf16_pair_w(tmem_base, a1, b1, idesc, accum);
        }
      }
      umma_commit_pair_w(empty(s), 3);
      tr.mark(0x20);
    }
    umma_commit_pair_w(acc_full, 3);
  }
}
__device__ __forceinline__ void xg_pair_finish(const XgTile& t, XgTracer& tr) {
  tr.mark(0x50);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                            // neither CTA leaves while its partner may still signal it
  tr.mark(0x60);
  if (t.warp == 8) tmem_dealloc_pair<256>(t.tmem_base);
}

template <bool A_UNIT_K, bool B_UNIT_K, int PARTS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kXgThreads, kXgMinBlocks) xgemm_pair_kernel(const GemmArgs a, const XgOperands o) {
  extern __shared__ uint8_t xg_smem[];
  XgTile t;
  XgTracer tr;
  xg_pair_mainloop<A_UNIT_K, B_UNIT_K, PARTS>(o, xg_smem, t, tr);
  xg_epilogue_gemm<1>(a, t, 4, blockIdx.y == 0);
  xg_pair_finish(t, tr);
}

// ---- t3d_linear_f32 (same contract as linear_f32_kernel, simt_ops.cuh): Y = act(X.W + bias + gbias[row / rows_per_group])
// * rowmask, optional max over the rows of each group.
template <int PARTS>
__device__ __forceinline__ void xg_epilogue_linear(const LinearArgs& a, const XgTile& t, int nchunks = 2) {
  if (t.warp < 8 && !(XG_DBG & 16)) {
    const int gm = t.m0 + (t.warp & 3) * 32 + t.lane;
    const bool row_ok = gm < a.M;
    const int g = (a.rows_per_group > 0 && row_ok) ? gm / a.rows_per_group : 0;
    const float rm = (a.rowmask && row_ok) ? a.rowmask[gm] : 1.0f;
    // the tile's 128 rows belong to one group: reduce over the warp's 32 rows in registers before the atomics
    const bool one_group = a.gmax && (a.rows_per_group % kXgBM == 0) && (t.m0 + kXgBM <= a.M);
#pragma unroll 1
    for (int ch = 0; ch < nchunks; ++ch) {
      const int c0 = (t.warp >> 2) * (32 * nchunks) + ch * 32;
      float v[32];
      xg_acc32<PARTS>(t, c0, v);
      // bias, then the per-group bias: 32 consecutive floats each, fetched as eight independent 128-bit loads before the
      // first use.  (One guarded scalar load per element left 32 dependent L2 round trips per chunk in the persistent
      // kernel, whose loaders keep L1 cold: its epilogue took 31 k cycles per tile in the clock64 trace.)
      const int gn0 = t.n0 + c0;
      const bool full = gn0 + 32 <= a.N;
      auto add_vec32 = [&](const float* __restrict__ p) {
        if (full && (((uintptr_t)p) & 15) == 0) {
          float4 q[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) q[i] = __ldg(reinterpret_cast<const float4*>(p) + i);
#pragma unroll
          for (int i = 0; i < 8; ++i) { v[4 * i] += q[i].x; v[4 * i + 1] += q[i].y; v[4 * i + 2] += q[i].z; v[4 * i + 3] += q[i].w; }
        } else {
          float q[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) q[j] = (gn0 + j < a.N) ? __ldg(p + j) : 0.0f;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += q[j];
        }
      };
      if (a.bias) add_vec32(a.bias + gn0);
      if (a.gbias && row_ok) add_vec32(a.gbias + (size_t)g * a.N + gn0);
      // the activation switch is hoisted out of the element loop: inside it the compiler evaluates every branch (tanhf
      // included) per element -- 28 k cycles per tile for the 4 epilogue warps of the persistent kernel (clock64 trace)
      if (a.act == ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (gn0 + j < a.N) ? fmaxf(v[j], 0.0f) * rm : 0.0f;
      } else if (a.act == ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (gn0 + j < a.N) ? v[j] * rm : 0.0f;
      } else {        // leaky ReLU / tanh (the refinement head); fully unrolled so that v[] stays in registers
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (gn0 + j < a.N) ? apply_act(v[j], a.act) * rm : 0.0f;
      }
      if (a.Y) xg_store_chunk<false>(t, a.Y, a.ldy, a.M, a.N, t.m0 + (t.warp & 3) * 32, t.n0 + c0, v);
      if (a.gmax) {
        if (one_group) {      // values are >= 0 here (ReLU / mask); lane c ends up with the max of column c0 + c
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) {
            const bool upper = (t.lane & off) != 0;
#pragma unroll
            for (int j = 0; j < off; ++j) {
              const float send = upper ? v[j] : v[j + off];
              const float keep = upper ? v[j + off] : v[j];
              v[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, off));
            }
          }
          const int gn = t.n0 + c0 + t.lane;
          if (gn < a.N) atomic_max_f32(a.gmax + (size_t)(t.m0 / a.rows_per_group) * a.N + gn, v[0]);
        } else if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int gn = t.n0 + c0 + j;
            if (gn < a.N) atomic_max_f32(a.gmax + (size_t)g * a.N + gn, v[j]);
          }
        }
      }
    }
  }
}

template <int PARTS>
__global__ void __launch_bounds__(kXgThreads, kXgMinBlocks) xlinear_kernel(const LinearArgs a, const XgOperands o) {
  extern __shared__ uint8_t xg_smem[];
  XgTile t;
  xg_mainloop<true, false, PARTS>(o, xg_smem, t);
  xg_epilogue_linear<PARTS>(a, t);
  xg_finish(t);
}
template <int PARTS>
__global__ void __launch_bounds__(kXgThreads, kXgMinBlocks) xlinear_pre_kernel(const LinearArgs a, const XgOperands o) {
  extern __shared__ uint8_t xg_smem[];
  XgTile t;
  xg_mainloop<true, false, PARTS, true>(o, xg_smem, t);
  xg_epilogue_linear<PARTS>(a, t);
  xg_finish(t);
}
// CTA-pair variant of t3d_linear_f32 (X k-contiguous, W [K, N]) for the engines with at most two operand pieces
template <int PARTS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kXgThreads, kXgMinBlocks) xlinear_pair_kernel(const LinearArgs a, const XgOperands o) {
  extern __shared__ uint8_t xg_smem[];
  XgTile t;
  XgTracer tr;
  xg_pair_mainloop<true, false, PARTS>(o, xg_smem, t, tr);
  xg_epilogue_linear<1>(a, t, 4);
  xg_pair_finish(t, tr);
}
// ---- persistent variant for forward / dgrad (k-contiguous A, pre-split B) ------------------------------------------------
// One CTA per SM walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  The per-tile costs of the one-tile kernel that the
// ablation exposed -- the epilogue (20 - 47 % of the forward layers) and the CTA skeleton (launch, TMEM allocation, barrier
// set-up: ~17 %) -- are taken off the critical path:
//   warps 0-7   loaders: A split (16 values per thread and stage), B stage by cp.async issued one stage ahead; 4-slot ring
//   warps 8-11  epilogue: drain accumulator set (tile & 1) while the MMAs of the next tile fill the other set
//   warp  12    MMA issue
// Shared memory: 4 stages x 6 images x 8 KB = 192 KB + 4 epilogue patches; TMEM: 2 x (main + small) = 512 columns.
#ifndef XG_PP_LOADERS
#define XG_PP_LOADERS 8
#endif
constexpr int kXgPPLoaders = XG_PP_LOADERS;                     // loader warps: 8 (16 rows each) or 16 (8 rows each)
constexpr int kXgPPBAhead = 2;                                  // stages the B copies run ahead of the A split (<= 2 with 4 slots)
constexpr int kXgPPVals = 128 / kXgPPLoaders;                   // A values per thread and stage (16 or 8)
constexpr int kXgPPSlots = 4;
// Geometry of the persistent kernels by number of operand pieces.  With three pieces the 12 operand images (192 KB) leave room
// for 4 epilogue warps (one staging patch each); with two or one the images are 128 / 64 KB and EIGHT epilogue warps drain an
// accumulator (lane quarter x column half) -- the 4-warp epilogue was the limit of these kernels (5.6 k cycles per 128 x 128
// tile against 3.3 k for its MMAs in the clock64 trace of the A-stationary kernel, profiles/r02_trace_xgemm_tc2.txt).
template <int PARTS> struct XgPP {
  static constexpr int EPI = PARTS == 3 ? 4 : 8;                                  // epilogue warps
  static constexpr int THREADS = (kXgPPLoaders + EPI + 1) * 32;                   // loaders + epilogue + the MMA warp
  static constexpr int MMA_WARP = kXgPPLoaders + EPI;
  static constexpr uint32_t B_IMG = 2u * PARTS * kXgImage;                         // A: PARTS x 2 blocks, then B the same
  static constexpr uint32_t PATCH = 4u * PARTS * kXgImage;                         // EPI x kXgPatchBytes
  static constexpr uint32_t BARS = PATCH + EPI * kXgPatchBytes;
  static constexpr uint32_t PARAMS = BARS + 128;                                   // [a_scale 128][a_shift 128] fp32 (K <= 128 here)
  static constexpr uint32_t SMEM = BARS + 128 + 1024 + 1024;
};

// shared address of (part p, ring slot s) of an operand whose images start at `base`: block s >> 1 of image p; the
// stage is half s & 1 of that block
__device__ __forceinline__ uint32_t xg_pp_block(uint32_t base, int p, int s) { return base + (uint32_t)(p * 2 + (s >> 1)) * kXgImage; }

template <int PARTS, bool LINEAR>
__global__ void __launch_bounds__(XgPP<PARTS>::THREADS, 1) xg_pp_kernel(const GemmArgs ga, const LinearArgs la, const XgOperands o) {
  extern __shared__ uint8_t xg_smem[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)xg_smem + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  using G = XgPP<PARTS>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = sbase + G::BARS;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (4 + s); };
  auto acc_full = [&](int b) { return bar0 + 8u * (8 + b); };
  auto acc_empty = [&](int b) { return bar0 + 8u * (10 + b); };
  const uint32_t tmem_slot = bar0 + 96u;
  const int ntm = (o.M + kXgBM - 1) / kXgBM, ntiles = ntm * o.ntn;
  const int nst = (o.K + kXgBK - 1) / kXgBK;
  const uint32_t a_img = sbase, b_img = sbase + G::B_IMG;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kXgPPSlots; ++s) { mbar_init(full(s), kXgPPLoaders); mbar_init(empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), G::EPI); }
    fence_barrier_init();
  }
  if (warp == G::MMA_WARP) tmem_alloc<512>(tmem_slot);
  if (o.a_scale != nullptr) {               // lazy BN tables of the A operand (K <= 128)
    float* tab = reinterpret_cast<float*>(smem + G::PARAMS);
    for (int i = threadIdx.x; i < o.K; i += G::THREADS) { tab[i] = __ldg(o.a_scale + i); tab[128 + i] = __ldg(o.a_shift + i); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + G::BARS + 96);
  XgTracer tr;      // roles: 0 loader warp 0, 1 MMA warp, 2 epilogue warp 0
  tr.init(o.trace, warp == 0 ? 0 : (warp == G::MMA_WARP ? 1 : 2), lane == 0 && (warp == 0 || warp == G::MMA_WARP || warp == kXgPPLoaders));
  tr.mark(0x02);

  if (warp < kXgPPLoaders) {
    // ------------------------------------------------------------------------------------------------ loaders
    const int rsub = lane >> 3, c = lane & 7, tid = threadIdx.x;
    constexpr int RW = 128 / kXgPPLoaders, NI = RW / 4;      // rows per warp, 128-bit loads per thread and stage
    auto rowof = [&](int i) { return warp * RW + (i >> 1) * 8 + (i & 1) * 2 + (rsub & 1) * 4 + (rsub >> 1); };
    const long long ld = o.lda, ld2 = 2 * ld, ld6 = 6 * ld;
    const int kofs = 4 * c;
    // work item g = (tile, stage), numbered consecutively over this CTA's tiles: slot g & 3, ring phase (g >> 2) & 1
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * nst;
    // Both streams (A loads, B copies) visit the work items in order: (tile, stage) advance by counters -- an integer
    // division is a ~100-cycle dependent chain, and three of them per call made up half of the stage time (clock64 trace).
    int l_it = 0, l_row0 = ((int)blockIdx.x / o.ntn) * kXgBM, l_tile = blockIdx.x;
    int c_it = 0, c_nt = (int)blockIdx.x % o.ntn, c_tile = blockIdx.x;
    auto load = [&](int g, float (&r)[kXgPPVals]) {
      const int it = l_it, row0 = l_row0, k0 = it * kXgBK;
      if (++l_it == nst) { l_it = 0; l_tile += gridDim.x; l_row0 = (l_tile / o.ntn) * kXgBM; }
      const int row = row0 + rowof(0);
      const float* q = o.A + (long long)row * ld + k0 + kofs;
      if (o.vecA && row0 + warp * RW + RW <= o.M && k0 + kXgBK <= o.K) {
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(q));
          r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
          q += (i & 1) ? ld6 : ld2;
        }
      } else {
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int ro = (i >> 1) * 8 + (i & 1) * 2;
#pragma unroll
          for (int e = 0; e < 4; ++e) r[4 * i + e] = (row + ro < o.M && k0 + kofs + e < o.K) ? __ldg(q + (long long)ro * ld + e) : 0.0f;
        }
      }
    };
    // B stage of work item g: PARTS x 128 rows x 4 chunks of 16 B from the pre-split workspace into slot g & 3
    auto copy_b = [&](int g) {
      const int it = c_it, nt = c_nt, s = g & 3;
      if (++c_it == nst) { c_it = 0; c_tile += gridDim.x; c_nt = c_tile % o.ntn; }
      mbar_wait_backoff(empty(s), ((g >> 2) & 1) ^ 1);
      const uint8_t* blk = o.bpre + ((size_t)nt * o.nkb + (it >> 1)) * kXgPreBlockBytes;
#pragma unroll
      for (int j = 0; j < PARTS * 512 / (kXgPPLoaders * 32); ++j) {
        const int id = tid + kXgPPLoaders * 32 * j, part = id >> 9, rem = id & 511;
        const uint32_t row = (uint32_t)(rem >> 2), ch = (uint32_t)(rem & 3);
        cp_async16(xg_pp_block(b_img, part, s) + sw128_offset(row, 4u * (s & 1) + ch),
                   blk + (size_t)part * kXgImage + sw128_offset(row, 4u * (it & 1) + ch));
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int e_it = 0;
    auto emit = [&](int g, float (&r)[kXgPPVals]) {
      const int s = g & 3;
      if (o.a_scale != nullptr) {      // lazy BN of A: the 4 channels of this thread in stage e_it
        const float4 sc = ld_shared_f4(sbase + G::PARAMS + 4u * (uint32_t)(e_it * kXgBK + kofs));
        const float4 sh = ld_shared_f4(sbase + G::PARAMS + 512u + 4u * (uint32_t)(e_it * kXgBK + kofs));
#pragma unroll
        for (int i = 0; i < NI; ++i) xg_bn4(&r[4 * i], sc, sh);
      }
      if (++e_it == nst) e_it = 0;
      // B runs kXgPPBAhead stages ahead (an L2 round trip is ~2 k cycles, longer than a stage): one commit group per stage
      if (g + kXgPPBAhead < total) copy_b(g + kXgPPBAhead);
      else asm volatile("cp.async.commit_group;" ::: "memory");
      tr.mark(0x18);
      // slot s is free: copy_b(g) waited for it kXgPPBAhead steps earlier (or in the prologue below)
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const uint32_t off = sw128_offset((uint32_t)rowof(i), (uint32_t)(4 * (s & 1) + (c >> 1))) + (uint32_t)(c & 1) * 8u;
        uint32_t w[3][2];
        xg_split4<PARTS>(&r[4 * i], w);
#pragma unroll
        for (int p = 0; p < PARTS; ++p) st_shared_v2(xg_pp_block(a_img, p, s) + off, w[p][0], w[p][1]);
      }
      tr.mark(0x28);
      asm volatile("cp.async.wait_group %0;" ::"n"(kXgPPBAhead) : "memory");      // all but the newest groups: this stage's B has landed
      tr.mark(0x2c);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full(s));
      tr.mark(0x30);
    };
    if (total > 0) {
      float r[kXgPreSets][kXgPPVals];
#pragma unroll
      for (int j = 0; j < kXgPreSets - 1; ++j) if (j < total) load(j, r[j]);
#pragma unroll
      for (int j = 0; j < kXgPPBAhead; ++j) {
        if (j < total) copy_b(j);
        else asm volatile("cp.async.commit_group;" ::: "memory");
      }
      for (int g0 = 0; g0 < total; g0 += kXgPreSets) {
#pragma unroll
        for (int j = 0; j < kXgPreSets; ++j) {
          const int g = g0 + j;
          if (g < total) {
            if (g + kXgPreSets - 1 < total) load(g + kXgPreSets - 1, r[(j + kXgPreSets - 1) % kXgPreSets]);
            tr.mark(0x10);
            emit(g, r[j]);
          }
        }
      }
    }
  } else if (warp == G::MMA_WARP) {
    // ------------------------------------------------------------------------------------------------ MMA issue
    int g = 0, lt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
      const int n0 = (tile % o.ntn) * kXgBN, ab = lt & 1;
      const int ncols = min(kXgBN, ((o.N - n0) + 15) & ~15);
      const uint32_t idesc = make_idesc_bf16(128, ncols);
      const uint32_t d_main = tmem_base + (uint32_t)ab * 256u, d_small = d_main + 128u;
      mbar_wait_w(acc_empty(ab), ((lt >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int it = 0; it < nst; ++it, ++g) {
        const int s = g & 3;
        mbar_wait_w(full(s), (g >> 2) & 1);
        tc_fence_after();
        tr.mark(0x10);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint32_t off = (uint32_t)(2 * (s & 1) + ks) * 32u;
          const uint64_t a1 = make_sdesc_k128(xg_pp_block(a_img, 0, s) + off), a2 = make_sdesc_k128(xg_pp_block(a_img, 1, s) + off),
                         a3 = make_sdesc_k128(xg_pp_block(a_img, 2, s) + off);
          const uint64_t b1 = make_sdesc_k128(xg_pp_block(b_img, 0, s) + off), b2 = make_sdesc_k128(xg_pp_block(b_img, 1, s) + off),
                         b3 = make_sdesc_k128(xg_pp_block(b_img, 2, s) + off);
          const uint32_t first = (it | ks) != 0;
          xg_issue<PARTS>(d_main, d_small, a1, a2, a3, b1, b2, b3, idesc, first);
        }
        umma_commit_w(empty(s));
        tr.mark(0x20);
      }
      umma_commit_w(acc_full(ab));
    }
  } else {
    // ------------------------------------------------------------------------------------------------ epilogue (warps 8-11)
    int lt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
      const int ab = lt & 1;
      XgTile t;
      t.sbase = sbase + G::PATCH; t.tmem_base = tmem_base + (uint32_t)ab * 256u;
      t.m0 = (tile / o.ntn) * kXgBM; t.n0 = (tile % o.ntn) * kXgBN; t.warp = warp - kXgPPLoaders; t.lane = lane;
      mbar_wait_backoff(acc_full(ab), (lt >> 1) & 1);
      tc_fence_after();
      tr.mark(0x40);
      if (LINEAR) xg_epilogue_linear<PARTS>(la, t, 16 / G::EPI);
      else xg_epilogue_gemm<PARTS>(ga, t, 16 / G::EPI, true);
      tc_fence_before();                                  // the tcgen05.ld of this tile are complete (wait::ld inside)
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(ab));
      tr.mark(0x50);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == G::MMA_WARP) tmem_dealloc<512>(tmem_base);
}
// ---- A-stationary variant for K <= 128 and several column tiles -----------------------------------------------------------
// Same roles and shared-memory map as xg_pp_kernel, but a CTA owns ROW tiles: the split images of its A block (at most 4
// stages) are written once and stay resident while the CTA walks all N / 128 column tiles, streaming only the pre-split B
// stages through the 4-slot ring.  A is loaded and split once per row tile instead of once per (row, column) tile, and the
// per-tile skeleton of the one-tile kernel is paid once per row tile.  The next row tile's A values are already in
// registers while the current one is multiplied; the images are rewritten once the last MMA of the row tile has retired
// (a_free), which costs a short bubble per row tile.
template <int PARTS, bool LINEAR>
__global__ void __launch_bounds__(XgPP<PARTS>::THREADS, 1) xg_as_kernel(const GemmArgs ga, const LinearArgs la, const XgOperands o) {
  extern __shared__ uint8_t xg_smem[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)xg_smem + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  using G = XgPP<PARTS>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = sbase + G::BARS;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (4 + s); };
  auto acc_full = [&](int b) { return bar0 + 8u * (8 + b); };
  auto acc_empty = [&](int b) { return bar0 + 8u * (10 + b); };
  const uint32_t a_ready = bar0 + 8u * 13, a_free = bar0 + 8u * 14;
  const uint32_t tmem_slot = bar0 + 96u;
  const int ntm = (o.M + kXgBM - 1) / kXgBM, ntn = o.ntn;
  const int nst = (o.K + kXgBK - 1) / kXgBK;                       // <= 4
  const uint32_t a_img = sbase, b_img = sbase + G::B_IMG;
  const int my_mt = (ntm - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // row tiles of this CTA

  if (threadIdx.x == 0) {
    for (int s = 0; s < kXgPPSlots; ++s) { mbar_init(full(s), kXgPPLoaders); mbar_init(empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), G::EPI); }
    mbar_init(a_ready, kXgPPLoaders); mbar_init(a_free, 1);
    fence_barrier_init();
  }
  if (warp == G::MMA_WARP) tmem_alloc<512>(tmem_slot);
  if (o.a_scale != nullptr) {               // lazy BN tables of the A operand (K <= 128)
    float* tab = reinterpret_cast<float*>(smem + G::PARAMS);
    for (int i = threadIdx.x; i < o.K; i += G::THREADS) { tab[i] = __ldg(o.a_scale + i); tab[128 + i] = __ldg(o.a_shift + i); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + G::BARS + 96);
  XgTracer tr;      // roles: 0 loader warp 0, 1 MMA warp, 2 epilogue warp 0
  tr.init(o.trace, warp == 0 ? 0 : (warp == G::MMA_WARP ? 1 : 2), lane == 0 && (warp == 0 || warp == G::MMA_WARP || warp == kXgPPLoaders));
  tr.mark(0x02);

  if (warp < kXgPPLoaders) {
    // ------------------------------------------------------------------------------------------------ loaders
    const int rsub = lane >> 3, c = lane & 7, tid = threadIdx.x;
    constexpr int RW = 128 / kXgPPLoaders, NI = RW / 4;
    auto rowof = [&](int i) { return warp * RW + (i >> 1) * 8 + (i & 1) * 2 + (rsub & 1) * 4 + (rsub >> 1); };
    const long long ld = o.lda, ld2 = 2 * ld, ld6 = 6 * ld;
    const int kofs = 4 * c;
    float rA[4][kXgPPVals];
    auto load_a = [&](int mt) {                                     // all stages of row tile mt -> registers
      const int row0 = mt * kXgBM, row = row0 + rowof(0);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        if (it < nst) {
          const int k0 = it * kXgBK;
          const float* q = o.A + (long long)row * ld + k0 + kofs;
          if (o.vecA && row0 + warp * RW + RW <= o.M && k0 + kXgBK <= o.K) {
#pragma unroll
            for (int i = 0; i < NI; ++i) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(q));
              rA[it][4 * i] = v.x; rA[it][4 * i + 1] = v.y; rA[it][4 * i + 2] = v.z; rA[it][4 * i + 3] = v.w;
              q += (i & 1) ? ld6 : ld2;
            }
          } else {
#pragma unroll
            for (int i = 0; i < NI; ++i) {
              const int ro = (i >> 1) * 8 + (i & 1) * 2;
#pragma unroll
              for (int e = 0; e < 4; ++e)
                rA[it][4 * i + e] = (row + ro < o.M && k0 + kofs + e < o.K) ? __ldg(q + (long long)ro * ld + e) : 0.0f;
            }
          }
        }
      }
    };
    auto store_a = [&]() {                                          // registers -> resident images, stage it in slot it
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        if (it < nst) {
          if (o.a_scale != nullptr) {      // lazy BN of A
            const float4 sc = ld_shared_f4(sbase + G::PARAMS + 4u * (uint32_t)(it * kXgBK + kofs));
            const float4 sh = ld_shared_f4(sbase + G::PARAMS + 512u + 4u * (uint32_t)(it * kXgBK + kofs));
#pragma unroll
            for (int i = 0; i < NI; ++i) xg_bn4(&rA[it][4 * i], sc, sh);
          }
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            const uint32_t off = sw128_offset((uint32_t)rowof(i), (uint32_t)(4 * (it & 1) + (c >> 1))) + (uint32_t)(c & 1) * 8u;
            const float* r = &rA[it][4 * i];
            uint32_t w[3][2];
            xg_split4<PARTS>(r, w);
#pragma unroll
            for (int p = 0; p < PARTS; ++p) st_shared_v2(xg_pp_block(a_img, p, it) + off, w[p][0], w[p][1]);
          }
        }
      }
    };
    // B stream: work item g = ((row tile, column tile), stage) in order; slot g & 3
    const int per_mt = ntn * nst, total = my_mt * per_mt;
    int c_it = 0, c_nt = 0;
    auto copy_b = [&](int g) {
      const int it = c_it, nt = c_nt, s = g & 3;
      if (++c_it == nst) { c_it = 0; if (++c_nt == ntn) c_nt = 0; }
      mbar_wait_backoff(empty(s), ((g >> 2) & 1) ^ 1);
      const uint8_t* blk = o.bpre + ((size_t)nt * o.nkb + (it >> 1)) * kXgPreBlockBytes;
#pragma unroll
      for (int j = 0; j < PARTS * 512 / (kXgPPLoaders * 32); ++j) {
        const int id = tid + kXgPPLoaders * 32 * j, part = id >> 9, rem = id & 511;
        const uint32_t row = (uint32_t)(rem >> 2), ch = (uint32_t)(rem & 3);
        cp_async16(xg_pp_block(b_img, part, s) + sw128_offset(row, 4u * (s & 1) + ch),
                   blk + (size_t)part * kXgImage + sw128_offset(row, 4u * (it & 1) + ch));
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (my_mt > 0) {
      load_a(blockIdx.x);
#pragma unroll
      for (int j = 0; j < kXgPPBAhead; ++j) {
        if (j < total) copy_b(j);
        else asm volatile("cp.async.commit_group;" ::: "memory");
      }
      int g = 0;
      for (int i = 0; i < my_mt; ++i) {
        tr.mark(0x10);
        mbar_wait_backoff(a_free, (i & 1) ^ 1);                     // every MMA of the previous row tile has retired
        tr.mark(0x11);
        store_a();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
        tr.mark(0x12);
        if (i + 1 < my_mt) load_a((int)blockIdx.x + (i + 1) * (int)gridDim.x);      // in flight during this row tile's B stream
        for (int q = 0; q < per_mt; ++q, ++g) {
          if (g + kXgPPBAhead < total) copy_b(g + kXgPPBAhead);
          else asm volatile("cp.async.commit_group;" ::: "memory");
          asm volatile("cp.async.wait_group %0;" ::"n"(kXgPPBAhead) : "memory");
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(full(g & 3));
          if ((q & 3) == 3) tr.mark(0x20);                           // B stages of one column tile delivered (nst = 4)
        }
      }
    }
  } else if (warp == G::MMA_WARP) {
    // ------------------------------------------------------------------------------------------------ MMA issue
    int g = 0, lt = 0;
    for (int i = 0; i < my_mt; ++i) {
      mbar_wait_w(a_ready, i & 1);
      tc_fence_after();
      for (int nt = 0; nt < ntn; ++nt, ++lt) {
        const int n0 = nt * kXgBN, ab = lt & 1;
        const int ncols = min(kXgBN, ((o.N - n0) + 15) & ~15);
        const uint32_t idesc = make_idesc_bf16(128, ncols);
        const uint32_t d_main = tmem_base + (uint32_t)ab * 256u, d_small = d_main + 128u;
        mbar_wait_w(acc_empty(ab), ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        tr.mark(0x10);
        for (int it = 0; it < nst; ++it, ++g) {
          const int s = g & 3;
          mbar_wait_w(full(s), (g >> 2) & 1);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t offa = (uint32_t)(2 * (it & 1) + ks) * 32u, offb = (uint32_t)(2 * (s & 1) + ks) * 32u;
            const uint64_t a1 = make_sdesc_k128(xg_pp_block(a_img, 0, it) + offa), a2 = make_sdesc_k128(xg_pp_block(a_img, 1, it) + offa),
                           a3 = make_sdesc_k128(xg_pp_block(a_img, 2, it) + offa);
            const uint64_t b1 = make_sdesc_k128(xg_pp_block(b_img, 0, s) + offb), b2 = make_sdesc_k128(xg_pp_block(b_img, 1, s) + offb),
                           b3 = make_sdesc_k128(xg_pp_block(b_img, 2, s) + offb);
            const uint32_t accum = (it | ks) != 0;
            xg_issue<PARTS>(d_main, d_small, a1, a2, a3, b1, b2, b3, idesc, accum);
          }
          umma_commit_w(empty(s));
        }
        umma_commit_w(acc_full(ab));
        tr.mark(0x20);
      }
      umma_commit_w(a_free);                                        // the resident A images may be rewritten
    }
  } else {
    // ------------------------------------------------------------------------------------------------ epilogue (4 warps)
    int lt = 0;
    for (int i = 0; i < my_mt; ++i) {
      const int mt = (int)blockIdx.x + i * (int)gridDim.x;
      for (int nt = 0; nt < ntn; ++nt, ++lt) {
        const int ab = lt & 1;
        XgTile t;
        t.sbase = sbase + G::PATCH; t.tmem_base = tmem_base + (uint32_t)ab * 256u;
        t.m0 = mt * kXgBM; t.n0 = nt * kXgBN; t.warp = warp - kXgPPLoaders; t.lane = lane;
        mbar_wait_backoff(acc_full(ab), (lt >> 1) & 1);
        tc_fence_after();
        tr.mark(0x40);
        if (LINEAR) xg_epilogue_linear<PARTS>(la, t, 16 / G::EPI);
        else xg_epilogue_gemm<PARTS>(ga, t, 16 / G::EPI, true);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(ab));
        tr.mark(0x50);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == G::MMA_WARP) tmem_dealloc<512>(tmem_base);
}
#endif

inline bool xg_aligned16(const void* p, long long ld) { return (((uintptr_t)p & 15) == 0) && (ld % 4 == 0); }

}  // namespace t3d
