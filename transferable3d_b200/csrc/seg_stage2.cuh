// Instance-segmentation stage 2 on tcgen05/TMEM: point_feat(64) -> conv6'(512) -> conv7(256) ->
// conv8(128) -> conv9(128) -> conv10(2) = mask logits (semisup_models.py:107-135, eval mode, BN folded).
//
// The reference tiles the 1024(+10)-wide global feature to every point and runs conv6 on the
// 1088-wide concat; here the global half of conv6 is folded into a per-frustum bias
// gbias[b] = b6 + [gfeat_b, one_hot_b] . W6[64:], so conv6' is a K=64 GEMM.
//
// Per 128-point tile (points on the UMMA M dimension, channels on N):
//   conv6' is produced in 4 blocks of 128 channels; each block's epilogue (+gbias, ReLU, bf16)
//   becomes a K=128 slice of conv7's A operand, so the 512-wide activation only ever exists as two
//   32 KB slices.  conv7 accumulates its 256 outputs in TMEM across the 4 slices.  conv10 (128->2) is
//   evaluated on CUDA cores from the fp32 conv9 epilogue registers.
// The two CTAs of a cluster share the weight stream (each fetches half of every 16 KB chunk and
// multicasts it); the input tile (stage 1's swizzled point_feat image) and gbias arrive by bulk copy;
// 8 epilogue warps split the accumulator columns.
// TMEM: R6[0]=cols 0..127, R6[1]=128..255 (conv6' blocks, later conv8 / conv9), R7=256..511 (conv7).
#pragma once
#include "common.cuh"
#include "chain_max.cuh"

namespace t3d {

constexpr int kSeg2Chunks = 26;   // per tile: 4 (W6') + 16 (W7) + 4 (W8) + 2 (W9), consumption order below
// arena = [26 chunk images][b7 256][b8 128][b9 128][W10 128x2][b10 2] fp32
constexpr int kSeg2Floats = 256 + 128 + 128 + 256 + 2;
constexpr size_t kSeg2ArenaBytes = (size_t)kSeg2Chunks * kChunkBytes + sizeof(float) * kSeg2Floats;

struct Seg2Args {
  const __nv_bfloat16* point_feat;   // stage-1 emit: per 256-point tile a [256 x 64] bf16 K-major SW128 image
  const float* gbias;                // [B, 512] fp32 per-frustum conv6 bias (global half + b6, BN folded)
  const uint8_t* arena;
  float* logits;                     // [B, N, 2]
  int B, N;
  unsigned long long* trace;
};

struct Seg2Smem {
  static constexpr int IN = 0;                       // [128 x 64] bf16, 16 KB
  static constexpr int A6 = 16384;                   // 2 x [128 x 128] bf16 (2 K-blocks each), 64 KB
  static constexpr int A7 = A6 + 65536;              // [128 x 256] bf16 (4 K-blocks), 64 KB
  static constexpr int RING = A7 + 65536;            // 4 x 16 KB
  static constexpr int GB = RING + kRingStages * kChunkBytes;   // 2 x 512 fp32
  static constexpr int FL = GB + 2 * 512 * 4;        // b7,b8,b9,W10,b10
  static constexpr int LX = FL + ((kSeg2Floats * 4 + 15) / 16) * 16;   // [128][2] fp32 partial logits of column half 1
  static constexpr int BARS = LX + 128 * 8;
  // ring_full[4], ring_empty[4], acc_full[3], acc_empty[3], in_ready, in_free, a6_ready[2], a6_free[2], a7_ready, a8_ready
  static constexpr int NBARS = 2 * kRingStages + 6 + 2 + 4 + 2;
  static constexpr int TMEM_SLOT = BARS + 8 * NBARS;
  static constexpr int TOTAL = TMEM_SLOT + 16;
};

constexpr int kSeg2Threads = 384;   // warp 0 weight producer, 1 MMA, 2 TMEM alloc, 3 input producer, 4-11 epilogue

__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kSeg2Threads, 1) seg_stage2_kernel(const Seg2Args args) {
  using L = Seg2Smem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  constexpr uint16_t kAllCtas = (1u << kClusterSize) - 1;

  const uint32_t bar0 = sbase + L::BARS;
  auto ring_full = [&](int s) { return bar0 + 8u * s; };
  auto ring_empty = [&](int s) { return bar0 + 8u * (kRingStages + s); };
  auto acc_full = [&](int r) { return bar0 + 8u * (2 * kRingStages + r); };        // r: 0,1 = R6[r], 2 = R7
  auto acc_empty = [&](int r) { return bar0 + 8u * (2 * kRingStages + 3 + r); };
  const uint32_t in_ready = bar0 + 8u * (2 * kRingStages + 6);
  const uint32_t in_free = bar0 + 8u * (2 * kRingStages + 7);
  auto a6_ready = [&](int b) { return bar0 + 8u * (2 * kRingStages + 8 + b); };
  auto a6_free = [&](int b) { return bar0 + 8u * (2 * kRingStages + 10 + b); };
  const uint32_t a7_ready = bar0 + 8u * (2 * kRingStages + 12);
  const uint32_t a8_ready = bar0 + 8u * (2 * kRingStages + 13);
  auto region_col = [&](int r) -> uint32_t { return r == 2 ? 256u : (uint32_t)(r * 128); };

  const int tiles_per_frustum = (args.N + 127) / 128;
  const int tiles256_per_frustum = (args.N + 255) / 256;
  const int num_tiles = args.B * tiles_per_frustum;
  const int ncl = gridDim.x / kClusterSize, cl = blockIdx.x / kClusterSize;
  const int cbegin = (int)(((long long)num_tiles * cl) / ncl);
  const int cend = (int)(((long long)num_tiles * (cl + 1)) / ncl);
  const int iters = (cend - cbegin + kClusterSize - 1) / kClusterSize;
  auto tile_of = [&](int i) { return min(cbegin + i * kClusterSize + (int)crank, cend - 1); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRingStages; ++s) { mbar_init(ring_full(s), 1); mbar_init(ring_empty(s), kClusterSize); }
    for (int r = 0; r < 3; ++r) { mbar_init(acc_full(r), 1); mbar_init(acc_empty(r), 8); }
    mbar_init(in_ready, 1); mbar_init(in_free, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(a6_ready(b), 8); mbar_init(a6_free(b), 1); }
    mbar_init(a7_ready, 8); mbar_init(a8_ready, 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(sbase + L::TMEM_SLOT);
  {
    const float* fsrc = reinterpret_cast<const float*>(args.arena + (size_t)kSeg2Chunks * kChunkBytes);
    float* fdst = reinterpret_cast<float*>(smem + L::FL);
    for (int i = threadIdx.x; i < kSeg2Floats; i += blockDim.x) fdst[i] = fsrc[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_SLOT);

  if (warp == 0) {
    // ================================================================ weight producer (half of every chunk, multicast)
    if (lane == 0) {
      constexpr uint32_t kHalf = kChunkBytes / kClusterSize;
      uint32_t it = 0;
      for (int i = 0; i < iters; ++i)
        for (int c = 0; c < kSeg2Chunks; ++c, ++it) {
          const int s = it % kRingStages;
          mbar_wait(ring_empty(s), ((it / kRingStages) & 1) ^ 1);
          mbar_arrive_expect_tx(ring_full(s), kChunkBytes);
          bulk_g2s_mc(sbase + L::RING + s * kChunkBytes + crank * kHalf, args.arena + (size_t)c * kChunkBytes + crank * kHalf,
                      kHalf, ring_full(s), kAllCtas);
        }
    }
  } else if (warp == 3) {
    // ================================================================ input producer: point_feat tile + gbias of the frustum
    if (lane == 0) {
      for (int i = 0; i < iters; ++i) {
        const int t = tile_of(i);
        const int fr = t / tiles_per_frustum, j = t % tiles_per_frustum;
        if (i > 0) mbar_wait(in_free, (i - 1) & 1);
        mbar_arrive_expect_tx(in_ready, 16384 + 2048);
        const size_t row0 = ((size_t)fr * tiles256_per_frustum + (j >> 1)) * 256 + (j & 1) * 128;
        bulk_g2s(sbase + L::IN, reinterpret_cast<const uint8_t*>(args.point_feat) + row0 * 128, 16384, in_ready);
        bulk_g2s(sbase + L::GB + (i & 1) * 2048, args.gbias + (size_t)fr * 512, 2048, in_ready);
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      uint32_t it = 0, acc_cnt[3] = {0, 0, 0}, a6r_cnt[2] = {0, 0};
      const uint32_t idesc = make_idesc_bf16(128, 128);
      Tracer tr; tr.init(args.trace, 1);
      auto mma_chunk = [&](uint32_t a_addr, uint32_t d, bool acc_first) {
        const int s = it % kRingStages;
        mbar_wait(ring_full(s), (it / kRingStages) & 1);
        tc_fence_after();
        const uint32_t b_addr = sbase + L::RING + s * kChunkBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(d, make_sdesc_k128(a_addr + k * 32), make_sdesc_k128(b_addr + k * 32), idesc, (acc_first || k > 0) ? 1u : 0u);
        umma_commit_mc(ring_empty(s), kAllCtas);
        ++it;
      };
      auto job6 = [&](int nb) {
        const int r = nb & 1;
        mbar_wait(acc_empty(r), (acc_cnt[r] & 1) ^ 1);
        tc_fence_after();
        tr.mark(0x20 + nb);
        mma_chunk(sbase + L::IN, tmem_base + region_col(r), false);
        umma_commit(acc_full(r)); acc_cnt[r]++;
        tr.mark(0x28 + nb);
        if (nb == 3) umma_commit(in_free);
      };
      const uint32_t idesc256 = make_idesc_bf16(128, 256);
      // conv7 slice nb: K=128 (2 K-blocks); per K-block the 256 weight rows are two adjacent ring stages
      // (rows 0-127 | rows 128-255, even-aligned in the chunk stream) read by ONE N=256 instruction
      auto job7 = [&](int nb) {
        const int b = nb & 1;
        mbar_wait(a6_ready(b), a6r_cnt[b] & 1); a6r_cnt[b]++;
        if (nb == 0) mbar_wait(acc_empty(2), (acc_cnt[2] & 1) ^ 1);
        tr.mark(0x30 + nb);
        tc_fence_after();
        for (int kb = 0; kb < 2; ++kb) {
          const int s = it % kRingStages;                  // even
          mbar_wait(ring_full(s), (it / kRingStages) & 1);
          mbar_wait(ring_full(s + 1), ((it + 1) / kRingStages) & 1);
          tc_fence_after();
          const uint32_t a_addr = sbase + L::A6 + b * 32768 + kb * 16384;
          const uint32_t b_addr = sbase + L::RING + s * kChunkBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + region_col(2), make_sdesc_k128(a_addr + k * 32), make_sdesc_k128(b_addr + k * 32), idesc256,
                      (nb | kb | k) != 0 ? 1u : 0u);
          umma_commit_mc(ring_empty(s), kAllCtas);
          umma_commit_mc(ring_empty(s + 1), kAllCtas);
          it += 2;
        }
        umma_commit(a6_free(b));
        if (nb == 3) { umma_commit(acc_full(2)); acc_cnt[2]++; }
        tr.mark(0x38 + nb);
      };
      for (int i = 0; i < iters; ++i) {
        const uint32_t tpar = i & 1;
        mbar_wait(in_ready, tpar);
        tr.mark(0x10);
        tc_fence_after();
        job6(0); job6(1); job7(0); job6(2); job6(3); job7(1); job7(2); job7(3);
        // conv8: A = A7 (4 K-blocks), D = R6[0]
        mbar_wait(a7_ready, tpar);
        mbar_wait(acc_empty(0), (acc_cnt[0] & 1) ^ 1);
        tr.mark(0x40);
        tc_fence_after();
        for (int kb = 0; kb < 4; ++kb) mma_chunk(sbase + L::A7 + kb * 16384, tmem_base + region_col(0), kb != 0);
        umma_commit(acc_full(0)); acc_cnt[0]++;
        tr.mark(0x41);
        // conv9: A = A8 (in the A6[0] buffer, 2 K-blocks), D = R6[1]
        mbar_wait(a8_ready, tpar);
        mbar_wait(acc_empty(1), (acc_cnt[1] & 1) ^ 1);
        tr.mark(0x50);
        tc_fence_after();
        for (int kb = 0; kb < 2; ++kb) mma_chunk(sbase + L::A6 + kb * 16384, tmem_base + region_col(1), kb != 0);
        umma_commit(acc_full(1)); acc_cnt[1]++;
        tr.mark(0x51);
      }
    }
  } else if (warp >= 4) {
    // ================================================================ epilogue warps: lane quarter = warp&3, column half = (warp-4)>>2
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const uint32_t fl = sbase + L::FL;     // byte addresses of the fp32 constants in smem
    const uint32_t b7 = fl, b8 = fl + 4 * 256, b9 = fl + 4 * 384, w10 = fl + 4 * 512, b10 = fl + 4 * 768;
    const uint32_t lx = sbase + L::LX;
    uint32_t acc_cnt[3] = {0, 0, 0};
    Tracer tr; tr.init((warp == 4 && lane == 0) ? args.trace : nullptr, 2);

    // this warp's share [cbeg, cbeg+span) of region r: +bias, ReLU, bf16, stored as K-blocks of 64 at obuf
    auto epi_to_smem = [&](int r, int cbeg, int span, uint32_t bias, uint32_t obuf) {
#pragma unroll 1
      for (int c0 = cbeg; c0 < cbeg + span; c0 += 64) {
        uint32_t va[32], vb[32];
        tmem_ld32(tmem_base + lane_sel + region_col(r) + c0, va);
        tmem_ld32(tmem_base + lane_sel + region_col(r) + c0 + 32, vb);
        tmem_ld_wait();
        tr.mark(0x70);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint32_t (&v)[32] = g == 0 ? va : vb;
          const int cg = c0 + g * 32;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = ld_shared_f4(bias + 4u * (cg + 4 * j));
            pk[2 * j] = pack_bf16_relu(__uint_as_float(v[4 * j]) + b4.x, __uint_as_float(v[4 * j + 1]) + b4.y);
            pk[2 * j + 1] = pack_bf16_relu(__uint_as_float(v[4 * j + 2]) + b4.z, __uint_as_float(v[4 * j + 3]) + b4.w);
          }
          const int kb = cg >> 6, j0 = (cg & 63) >> 3;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            st_shared_v4(obuf + kb * 16384 + sw128_offset(row, j0 + jj), pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
        }
        tr.mark(0x71);
      }
    };
    auto release = [&](int r, uint32_t ready_bar) {
      tc_fence_before();
      fence_proxy_async_smem();
      tr.mark(0x72);
      __syncwarp();
      if (lane == 0) { if (ready_bar) mbar_arrive(ready_bar); mbar_arrive(acc_empty(r)); }
    };

    for (int i = 0; i < iters; ++i) {
      const int t = tile_of(i);
      const int fr = t / tiles_per_frustum;
      const int start = (t % tiles_per_frustum) * 128;
      const int npts = min(128, args.N - start);
      const uint32_t gb = sbase + L::GB + (i & 1) * 2048;
      mbar_wait(in_ready, i & 1);                    // gbias of this tile has landed in smem
      tr.mark(0x10);
      for (int nb = 0; nb < 4; ++nb) {
        const int r = nb & 1;
        mbar_wait(acc_full(r), acc_cnt[r] & 1); acc_cnt[r]++;
        if (nb >= 2) mbar_wait(a6_free(r), 0);       // commit #(2*tile) of this buffer (two commits per tile)
        tr.mark(0x20 + nb);
        tc_fence_after();
        epi_to_smem(r, half * 64, 64, gb + 4u * (nb * 128), sbase + L::A6 + r * 32768);
        release(r, a6_ready(r));
        tr.mark(0x28 + nb);
      }
      mbar_wait(acc_full(2), acc_cnt[2] & 1); acc_cnt[2]++;
      tr.mark(0x30);
      tc_fence_after();
      epi_to_smem(2, half * 128, 128, b7, sbase + L::A7);
      release(2, a7_ready);
      tr.mark(0x31);
      mbar_wait(acc_full(0), acc_cnt[0] & 1); acc_cnt[0]++;
      tr.mark(0x40);
      tc_fence_after();
      epi_to_smem(0, half * 64, 64, b8, sbase + L::A6);          // act8 reuses the A6[0] buffer
      release(0, a8_ready);
      tr.mark(0x41);
      mbar_wait(acc_full(1), acc_cnt[1] & 1); acc_cnt[1]++;
      tr.mark(0x50);
      tc_fence_after();
      // conv9 epilogue + conv10 (128 -> 2) in fp32: each column half reduces its 64 channels
      float l0 = 0.f, l1 = 0.f;
      {
        uint32_t va[32], vb[32];
        const int c0 = half * 64;
        tmem_ld32(tmem_base + lane_sel + region_col(1) + c0, va);
        tmem_ld32(tmem_base + lane_sel + region_col(1) + c0 + 32, vb);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const uint32_t (&v)[32] = g == 0 ? va : vb;
          const int cg = c0 + g * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bb = ld_shared_f4(b9 + 4u * (cg + j));
            const float4 w01 = ld_shared_f4(w10 + 8u * (cg + j)), w23 = ld_shared_f4(w10 + 8u * (cg + j + 2));
            const float a0 = fmaxf(__uint_as_float(v[j]) + bb.x, 0.0f), a1 = fmaxf(__uint_as_float(v[j + 1]) + bb.y, 0.0f);
            const float a2 = fmaxf(__uint_as_float(v[j + 2]) + bb.z, 0.0f), a3 = fmaxf(__uint_as_float(v[j + 3]) + bb.w, 0.0f);
            l0 = fmaf(a0, w01.x, l0); l1 = fmaf(a0, w01.y, l1);
            l0 = fmaf(a1, w01.z, l0); l1 = fmaf(a1, w01.w, l1);
            l0 = fmaf(a2, w23.x, l0); l1 = fmaf(a2, w23.y, l1);
            l0 = fmaf(a3, w23.z, l0); l1 = fmaf(a3, w23.w, l1);
          }
        }
      }
      if (half == 1) st_shared_f2(lx + 8u * row, l0, l1);
      release(1, 0);
      named_bar_sync(1 + q, 64);                     // the two warps of this lane quarter
      if (half == 0) {
        const float2 o = ld_shared_f2(lx + 8u * row), bz = ld_shared_f2(b10);
        if (row < npts)
          *reinterpret_cast<float2*>(args.logits + ((size_t)fr * args.N + start + row) * 2) =
              make_float2(l0 + o.x + bz.x, l1 + o.y + bz.y);
      }
      named_bar_sync(1 + q, 64);                     // lx may be overwritten by the next tile only after it was read
      tr.mark(0x51);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

}  // namespace t3d
