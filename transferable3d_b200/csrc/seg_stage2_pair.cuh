// Instance-segmentation stage 2 on tcgen05/TMEM as a CTA pair (cta_group::2): point_feat(64) -> conv6'(512) -> conv7(256) ->
// conv8(128) -> conv9(128) -> conv10(2) = mask logits (sunrgbd_detection/semisup_models.py:107-135, eval mode, BN folded).
//
// The reference tiles the 1024(+10)-wide global feature to every point and runs conv6 on the 1088-wide concat; here the
// global half of conv6 is folded into a per-frustum bias gbias[b] = b6 + [gfeat_b, one_hot_b] . W6[64:], so conv6' is a
// K=64 GEMM.  Per 128-point tile (points on the UMMA M dimension, channels on N) conv6' is produced in 4
// blocks of 128 channels; each block's epilogue (+gbias, ReLU, bf16) becomes a K=128 slice of conv7's A operand, so the
// 512-wide activation never exists as a whole; conv7 accumulates its 256 outputs in TMEM across the 4 slices; conv10
// (128 -> 2) is evaluated on CUDA cores from the fp32 conv9 epilogue registers.
//
// Software-pipelined: TWO 128-point tiles are in flight per CTA:
//   phase 1 of tile i    conv6' blocks (K=64) -> e6 (+gbias, ReLU, bf16) -> conv7 accumulation (K=512 in 4 slices)
//   phase 2 of tile i-1  e7 -> conv8 -> e8 -> conv9 -> e9 + conv10 (fp32, CUDA cores) -> logits
// run concurrently: one MMA-issuer warp interleaves both phases in a fixed order, and two groups of 8 epilogue warps serve
// them (P1: the four e6 blocks, P2: e7 / e8 / e9).  TMEM: R6 = cols 0..127 (conv6' block), R89 = 128..255 (conv8, then
// conv9), R7 = 256..511 (conv7); accumulators are released to the MMA warp as soon as the epilogue has them in registers.
//
// Why a pair: the one-CTA version of this kernel (round 1 / first half of round 2: every CTA of the cluster ingested every
// weight chunk through a multicast ring) measured 10.8 k cycles per tile against a 6.7 k MMA floor; per tile its MMA operand
// reads (704 KB), weight ring writes (416 KB) and epilogue stores (224 KB) add up to 1.36 MB = 10.6 k cycles at the 128 B / clk
// of the shared-memory crossbar.  With cta_group::2 the two CTAs of a cluster form one UMMA of M = 256 (their two tiles) and
// the B operand (weights) is SPLIT between the two shared memories: every CTA holds, ingests and feeds to the tensor core
// only half of every weight chunk.  Per tile and SM: operand reads 496 KB, ring writes 176 KB, epilogue stores 224 KB.
// Measured: 5.37 -> 5.19 ms per 8192 frustums (DESIGN.md section 4 has the variants that were measured and not kept:
// conv7 accumulator halves drained separately, 3 / 4 activation slots, all-N=128 conv7 with the conv6' weights in the ring).
//
//   leader (cluster rank 0)  issues every tcgen05.mma / tcgen05.commit (multicast to the barriers of both CTAs); its MMA warp
//                            runs converged with one elected lane (common.cuh: issuing from a divergent region costs ~85
//                            cycles per instruction)
//   peer   (cluster rank 1)  its epilogue warps arrive REMOTELY on the leader's barriers (accumulator drained, operand
//                            written); two relay lanes forward "weight stage landed" / "input tile landed" to the leader
//
// Shared memory per CTA: input tile 16 KB, two conv6' K-block slots 32 KB, phase-2 buffer 64 KB (conv7 activation = A of
// conv8, then the 32 KB conv8 activation = A of conv9), this CTA's half of the conv6' weights RESIDENT (32 KB: those four
// chunks never go through the ring), weight ring 8 x 8 KB.  Ring stage = this CTA's 64 rows of an N = 128 chunk (conv8,
// conv9); a conv7 K-block takes two adjacent stages = one whole 128-row chunk per CTA (rows 0-127 in the leader, 128-255 in
// the peer) and is ONE N = 256 instruction per K step.  Stage counts per iteration are even everywhere, so the two stages of
// a conv7 K-block never straddle the ring's wrap.
#pragma once
#include "common.cuh"
#include "chain_max.cuh"

namespace t3d {

constexpr int kSeg2Chunks = 26;   // per tile: 4 (W6') + 16 (W7) + 4 (W8) + 2 (W9); arena order in t3d_pack_seg2
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

constexpr int kSeg2PThreads = 640;   // warp 0 weight producer, 1 MMA (leader) / ring relay (peer), 2 TMEM alloc (+ input relay in the peer),
                                     // 3 input producer, 4-11 P1 epilogue, 12-19 P2 epilogue

struct Seg2QSmem {
  static constexpr int STAGES = 8;
  static constexpr int STAGE_BYTES = kChunkBytes / 2;
  static constexpr int IN = 0;                        // [128 x 64] bf16 point_feat tile, 16 KB
  static constexpr int A6 = 16384;                    // 2 K-block slots [128 x 64] bf16 of the conv6' activation, 16 KB each
  static constexpr int P2 = A6 + 2 * 16384;           // phase-2 buffer: conv7 activation [128 x 256] (4 K-blocks), then conv8's [128 x 128]
  static constexpr int W6 = P2 + 65536;               // resident: rows 64 r .. 64 r + 63 of the four conv6' chunks, 4 x 8 KB
  static constexpr int RING = W6 + 4 * STAGE_BYTES;   // STAGES x STAGE_BYTES
  static constexpr int GB = RING + STAGES * STAGE_BYTES;   // 2 x 512 fp32 gbias
  static constexpr int FL = GB + 2 * 512 * 4;         // b7,b8,b9,W10,b10
  static constexpr int LX = FL + ((kSeg2Floats * 4 + 15) / 16) * 16;   // [128][2] fp32 partial logits of column half 1
  static constexpr int BARS = LX + 128 * 8;
  // ring_full[S] ring_empty[S] in_ready in_free r6_full r6_empty a6_ready[2] a6_free[2] r7_full r7_empty a7_ready a8_ready r89_full r89_empty w6_ready
  static constexpr int NBARS = 2 * STAGES + 15;
  static constexpr int TMEM_SLOT = BARS + 8 * NBARS;
  static constexpr int TOTAL = TMEM_SLOT + 16;
};
static_assert(Seg2QSmem::TOTAL + 1024 <= 232448, "seg_stage2_pair: shared memory budget");
static_assert(kClusterSize == 2, "seg_stage2_pair: one CTA pair per cluster");

__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kSeg2PThreads, 1) seg_stage2_pair_kernel(const Seg2Args args) {
  using L = Seg2QSmem;
  constexpr int kStages = L::STAGES;
  constexpr uint32_t kStageBytes = L::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  constexpr uint16_t kBoth = 3;

  const uint32_t bar0 = sbase + L::BARS;
  auto ring_full = [&](int s) { return bar0 + 8u * s; };
  auto ring_empty = [&](int s) { return bar0 + 8u * (kStages + s); };
  constexpr int B0 = 2 * kStages;
  const uint32_t in_ready = bar0 + 8u * (B0 + 0), in_free = bar0 + 8u * (B0 + 1);
  const uint32_t r6_full = bar0 + 8u * (B0 + 2), r6_empty = bar0 + 8u * (B0 + 3);
  auto a6_ready = [&](int b) { return bar0 + 8u * (B0 + 4 + b); };
  auto a6_free = [&](int b) { return bar0 + 8u * (B0 + 6 + b); };
  const uint32_t r7_full = bar0 + 8u * (B0 + 8), r7_empty = bar0 + 8u * (B0 + 9);
  const uint32_t a7_ready = bar0 + 8u * (B0 + 10), a8_ready = bar0 + 8u * (B0 + 11);
  const uint32_t r89_full = bar0 + 8u * (B0 + 12), r89_empty = bar0 + 8u * (B0 + 13);
  const uint32_t w6_ready = bar0 + 8u * (B0 + 14);
  constexpr uint32_t kR6 = 0, kR89 = 128, kR7 = 256;
  const int tiles_per_frustum = (args.N + 127) / 128;
  const int tiles256_per_frustum = (args.N + 255) / 256;
  const int num_tiles = args.B * tiles_per_frustum;
  const int ncl = gridDim.x / kClusterSize, cl = blockIdx.x / kClusterSize;
  const int cbegin = (int)(((long long)num_tiles * cl) / ncl);
  const int cend = (int)(((long long)num_tiles * (cl + 1)) / ncl);
  const int iters = (cend - cbegin + kClusterSize - 1) / kClusterSize;
  auto tile_of = [&](int i) { return min(cbegin + i * kClusterSize + (int)crank, cend - 1); };
  // epilogue -> MMA-warp signals: the MMA warp lives in the leader CTA
  auto arrive_mma = [&](uint32_t bar) { if (leader) mbar_arrive(bar); else mbar_arrive_remote(bar, 0); };

  if (threadIdx.x == 0) {
    // the leader's ring_full / in_ready also count the peer's forwarded completion; its epilogue-side barriers count the
    // 8 warps of both CTAs; everything the MMA warp signals is one multicast commit per CTA
    const uint32_t fwd = leader ? 2 : 1;
    for (int s = 0; s < kStages; ++s) { mbar_init(ring_full(s), fwd); mbar_init(ring_empty(s), 1); }
    mbar_init(in_ready, fwd); mbar_init(in_free, 1);
    mbar_init(r6_full, 1); mbar_init(r6_empty, 16);
    for (int b = 0; b < 2; ++b) { mbar_init(a6_ready(b), 16); mbar_init(a6_free(b), 1); }
    mbar_init(r7_full, 1); mbar_init(r7_empty, 16);
    mbar_init(a7_ready, 16); mbar_init(a8_ready, 16);
    mbar_init(r89_full, 1); mbar_init(r89_empty, 16);
    mbar_init(w6_ready, 1);
    fence_barrier_init();
    // resident conv6' weights: this CTA's 64 rows of arena chunks 0, 1, 6, 7 (c6(0) .. c6(3))
    mbar_arrive_expect_tx(w6_ready, 4 * kStageBytes);
    for (int nb = 0; nb < 4; ++nb) {
      const int c = nb < 2 ? nb : 4 + nb;
      bulk_g2s(sbase + L::W6 + nb * kStageBytes, args.arena + (size_t)c * kChunkBytes + crank * kStageBytes, kStageBytes, w6_ready);
    }
    mbar_wait(w6_ready, 0);
  }
  if (warp == 2) tmem_alloc_pair<512>(sbase + L::TMEM_SLOT);
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

  // ring stages consumed per iteration, in MMA order: c7(0) [4] | conv8 of tile i-1 [4] | c7(1) [4] | conv9 of tile i-1 [2] |
  // c7(2) [4] | c7(3) [4]; the phase-2 items are absent in the first iteration, the phase-1 items in the last
  if (warp == 0) {
    // ================================================================ weight producer: this CTA's half of every chunk, kept local
    if (lane == 0) {
      uint32_t it = 0;
      auto fill = [&](const uint8_t* src) {
        const int s = it % kStages;
        mbar_wait(ring_empty(s), ((it / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(ring_full(s), kStageBytes);
        bulk_g2s(sbase + L::RING + s * kStageBytes, src, kStageBytes, ring_full(s));
        ++it;
      };
      auto chunk = [&](int c) { return args.arena + (size_t)c * kChunkBytes; };
      // conv7 slice nb: per K-block this CTA's 128 output rows = the whole chunk (c7 base + 2 kb + rank), two stages
      auto w7 = [&](int nb) {
        const int base = nb == 0 ? 2 : 4 + 4 * nb;
        for (int kb = 0; kb < 2; ++kb) {
          const uint8_t* src = chunk(base + 2 * kb + (int)crank);
          fill(src); fill(src + kStageBytes);
        }
      };
      auto w_half = [&](int c) { fill(chunk(c) + crank * kStageBytes); };      // N = 128 chunk: rows 64 r .. 64 r + 63
      for (int i = 0; i <= iters; ++i) {
        const bool p1 = i < iters, p2 = i > 0;
        if (p1) w7(0);
        if (p2) for (int kb = 0; kb < 4; ++kb) w_half(20 + kb);
        if (p1) w7(1);
        if (p2) for (int kb = 0; kb < 2; ++kb) w_half(24 + kb);
        if (p1) { w7(2); w7(3); }
      }
    }
  } else if (!leader && warp == 1) {
    // ================================================================ peer: relay the ring's TMA completions to the leader's MMA warp
    if (lane == 0) {
      const uint32_t total = (uint32_t)iters * 22u;
      for (uint32_t it = 0; it < total; ++it) {
        const int s = it % kStages;
        mbar_wait(ring_full(s), (it / kStages) & 1);
        mbar_arrive_remote(ring_full(s), 0);
      }
    }
  } else if (!leader && warp == 2) {
    // ================================================================ peer: relay "input tile landed" to the leader's MMA warp
    if (lane == 0) {
      for (int i = 0; i < iters; ++i) {
        mbar_wait(in_ready, i & 1);
        mbar_arrive_remote(in_ready, 0);
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
    // ================================================================ MMA issuer (leader): one instruction drives both SMs; the whole
    // warp runs this loop converged, one elected lane issues (umma_*_pair_w); every value below is warp-uniform
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t it = 0, n_r6 = 0, n_r89 = 0, n_a60 = 0, n_a61 = 0;
    const uint32_t idesc128 = make_idesc_bf16(256, 128), idesc256 = make_idesc_bf16(256, 256);
    const uint32_t p2buf = sbase + L::P2;
    Tracer tr; tr.init(lane == 0 ? args.trace : nullptr, 1);
    // one N = 128 weight chunk (64 rows of it per CTA) x K = 64 from the ring: D (+)= A[256 x 64] . chunk^T
    auto mma_chunk = [&](uint32_t a_addr, uint32_t d, bool acc_first) {
      const uint32_t s = it % kStages;
      mbar_wait_w(ring_full(s), (it / kStages) & 1);
      tc_fence_after();
      const uint64_t ad = make_sdesc_k128(a_addr), bd = make_sdesc_k128(sbase + L::RING + s * kStageBytes);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_pair_w(d, ad + 2u * k, bd + 2u * k, idesc128, (acc_first || k > 0) ? 1u : 0u);   // +32 B per K step (>>4)
      umma_commit_pair_w(ring_empty(s), kBoth);
      ++it;
    };
    for (int i = 0; i <= iters; ++i) {
      const bool p1 = i < iters, p2 = i > 0;
      auto c6 = [&](int nb) {      // conv6' block nb from the resident weights
        mbar_wait_w(r6_empty, (n_r6 & 1) ^ 1);
        tc_fence_after();
        tr.mark(0x20 + nb);
        const uint64_t ad = make_sdesc_k128(sbase + L::IN), bd = make_sdesc_k128(sbase + L::W6 + nb * kStageBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_pair_w(tm + kR6, ad + 2u * k, bd + 2u * k, idesc128, k > 0 ? 1u : 0u);
        umma_commit_pair_w(r6_full, kBoth); n_r6++;
        if (nb == 3) umma_commit_pair_w(in_free, kBoth);
        tr.mark(0x28 + nb);
      };
      // conv7 slice nb = two K-blocks of 64 (the two A6 slots); per K-block the 256 weight rows are one chunk in each CTA
      // (two adjacent ring stages): one N = 256 instruction per K step
      auto c7 = [&](int nb) {
        tr.mark(0x30 + nb);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          if (kb == 0) { mbar_wait_w(a6_ready(0), n_a60 & 1); n_a60++; } else { mbar_wait_w(a6_ready(1), n_a61 & 1); n_a61++; }
          if (nb == 0 && kb == 0) mbar_wait_w(r7_empty, (i & 1) ^ 1);
          tr.mark(0x70 + kb);
          const uint32_t s = it % kStages;                 // even; s + 1 < kStages
          mbar_wait_w(ring_full(s), (it / kStages) & 1);
          mbar_wait_w(ring_full(s + 1), (it / kStages) & 1);
          tr.mark(0x72 + kb);
          tc_fence_after();
          const uint64_t ad = make_sdesc_k128(sbase + L::A6 + kb * 16384);
          const uint64_t bd = make_sdesc_k128(sbase + L::RING + s * kStageBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_pair_w(tm + kR7, ad + 2u * k, bd + 2u * k, idesc256, (nb | kb | k) != 0 ? 1u : 0u);
          umma_commit_pair_w(ring_empty(s), kBoth);
          umma_commit_pair_w(ring_empty(s + 1), kBoth);
          umma_commit_pair_w(a6_free(kb), kBoth);
          it += 2;
        }
        if (nb == 3) umma_commit_pair_w(r7_full, kBoth);
        tr.mark(0x38 + nb);
      };
      if (p1) { mbar_wait_w(in_ready, i & 1); tc_fence_after(); }
      tr.mark(0x10);
      if (p1) { c6(0); c6(1); c7(0); c6(2); }
      if (p2) {   // conv8 of tile i-1: A = conv7 activation (4 K-blocks) -> R89
        mbar_wait_w(a7_ready, (i - 1) & 1);
        mbar_wait_w(r89_empty, (n_r89 & 1) ^ 1);
        tr.mark(0x40);
        tc_fence_after();
        for (int kb = 0; kb < 4; ++kb) mma_chunk(p2buf + kb * 16384, tm + kR89, kb != 0);
        umma_commit_pair_w(r89_full, kBoth); n_r89++;
        tr.mark(0x41);
      }
      if (p1) { c7(1); c6(3); }
      if (p2) {   // conv9 of tile i-1: A = conv8 activation (2 K-blocks) -> R89
        mbar_wait_w(a8_ready, (i - 1) & 1);
        mbar_wait_w(r89_empty, (n_r89 & 1) ^ 1);
        tr.mark(0x50);
        tc_fence_after();
        for (int kb = 0; kb < 2; ++kb) mma_chunk(p2buf + kb * 16384, tm + kR89, kb != 0);
        umma_commit_pair_w(r89_full, kBoth); n_r89++;
        tr.mark(0x51);
      }
      if (p1) { c7(2); c7(3); }
    }
  } else if (warp >= 4) {
    // ================================================================ epilogue warps: lane quarter = warp&3, column half per group
    const bool is_p1 = warp < 12;
    const int q = warp & 3;
    const int half = ((warp - 4) & 7) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const uint32_t fl = sbase + L::FL;     // byte addresses of the fp32 constants in smem
    const uint32_t b7 = fl, b8 = fl + 4 * 256, b9 = fl + 4 * 384, w10 = fl + 4 * 512, b10 = fl + 4 * 768;
    const uint32_t lx = sbase + L::LX;
    Tracer tr; tr.init(((warp == 4 || warp == 12) && lane == 0) ? args.trace : nullptr, is_p1 ? 2 : 3);

    // 64 accumulator columns [c0, c0+64) of TMEM region `rcol` into registers
    auto load64 = [&](uint32_t rcol, int c0, uint32_t (&va)[32], uint32_t (&vb)[32]) {
      tmem_ld32(tmem_base + lane_sel + rcol + c0, va);
      tmem_ld32(tmem_base + lane_sel + rcol + c0 + 32, vb);
      tmem_ld_wait();
    };
    // the accumulator region is free again as soon as this warp's values are in registers
    auto release_acc = [&](uint32_t empty_bar) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_mma(empty_bar);
    };
    // 32 columns [cg, cg+32) of a layer's output: +bias, ReLU, bf16, stored into the K-major SW128 operand whose K-block
    // kb = cg / 64 lives at obuf + kb * 16 KB
    auto store32 = [&](const uint32_t (&v)[32], int cg, uint32_t bias, uint32_t obuf) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = ld_shared_f4(bias + 4u * (cg + 4 * j));
        float x0, x1, x2, x3;      // packed adds: the fma pipe of the four epilogue warps per scheduler is a limit of this kernel
        xg_upk2(xg_add2(xg_pk2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), xg_pk2(b4.x, b4.y)), x0, x1);
        xg_upk2(xg_add2(xg_pk2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), xg_pk2(b4.z, b4.w)), x2, x3);
        pk[2 * j] = pack_bf16_relu(x0, x1);
        pk[2 * j + 1] = pack_bf16_relu(x2, x3);
      }
      const int kb = cg >> 6, j0 = (cg & 63) >> 3;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        st_shared_v4(obuf + kb * 16384 + sw128_offset(row, j0 + jj), pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
    };
    auto store64 = [&](const uint32_t (&va)[32], const uint32_t (&vb)[32], int c0, uint32_t bias, uint32_t obuf) {
      store32(va, c0, bias, obuf);
      store32(vb, c0 + 32, bias, obuf);
    };
    auto publish = [&](uint32_t ready_bar) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) arrive_mma(ready_bar);
    };

    if (is_p1) {
      // ---------------------------------------------------------------- P1: the four conv6' blocks of every tile
      // each conv6' block of 128 channels leaves as two K-blocks of 64 (slot = K-block parity); this warp owns the
      // 32 columns [half*32, +32) of either K-block
      uint32_t n6 = 0, n_slot = 0;
      for (int i = 0; i < iters; ++i) {
        const uint32_t gb = sbase + L::GB + (i & 1) * 2048;
        mbar_wait(in_ready, i & 1);                    // gbias of this tile has landed in smem
        tr.mark(0x10);
        for (int nb = 0; nb < 4; ++nb, ++n_slot) {
          uint32_t va[32], vb[32];
          mbar_wait(r6_full, n6 & 1); n6++;
          tc_fence_after();
          tr.mark(0x20 + nb);
          tmem_ld32(tmem_base + lane_sel + kR6 + half * 32, va);
          tmem_ld32(tmem_base + lane_sel + kR6 + 64 + half * 32, vb);
          tmem_ld_wait();
          tr.mark(0x60);
          release_acc(r6_empty);
          // slot use number n_slot waits for the conv7 K-block that read use n_slot-1
          mbar_wait(a6_free(0), (n_slot & 1) ^ 1);
          tr.mark(0x61);
          store32(va, half * 32, gb + 4u * (nb * 128), sbase + L::A6);
          publish(a6_ready(0));
          tr.mark(0x62);
          mbar_wait(a6_free(1), (n_slot & 1) ^ 1);
          tr.mark(0x63);
          store32(vb, half * 32, gb + 4u * (nb * 128 + 64), sbase + L::A6 + 16384);
          publish(a6_ready(1));
          tr.mark(0x28 + nb);
        }
      }
    } else {
      // ---------------------------------------------------------------- P2: e7, e8, e9 + conv10 of every tile
      uint32_t n89 = 0;
      for (int i = 0; i < iters; ++i) {
        const int t = tile_of(i);
        const int fr = t / tiles_per_frustum;
        const int start = (t % tiles_per_frustum) * 128;
        const int npts = min(128, args.N - start);
        const uint32_t tb = sbase + L::P2;
        tr.mark(0x10);
        mbar_wait(r7_full, i & 1);
        tc_fence_after();
        tr.mark(0x30);
        {
          uint32_t va[32], vb[32];
          load64(kR7, half * 128, va, vb);
          store64(va, vb, half * 128, b7, tb);
          load64(kR7, half * 128 + 64, va, vb);
          release_acc(r7_empty);
          store64(va, vb, half * 128 + 64, b7, tb);
        }
        publish(a7_ready);
        tr.mark(0x31);
        mbar_wait(r89_full, n89 & 1); n89++;
        tc_fence_after();
        tr.mark(0x40);
        {
          uint32_t va[32], vb[32];
          load64(kR89, half * 64, va, vb);
          release_acc(r89_empty);
          store64(va, vb, half * 64, b8, tb);          // conv8 activation reuses the first 32 KB of the phase-2 buffer
        }
        publish(a8_ready);
        tr.mark(0x41);
        mbar_wait(r89_full, n89 & 1); n89++;
        tc_fence_after();
        tr.mark(0x50);
        // conv9 epilogue + conv10 (128 -> 2) in fp32: each column half reduces its 64 channels
        float l0 = 0.f, l1 = 0.f;
        unsigned long long l01 = 0ull;      // (l0, l1) accumulated with packed FMAs
        {
          uint32_t va[32], vb[32];
          const int c0 = half * 64;
          load64(kR89, c0, va, vb);
          release_acc(r89_empty);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint32_t (&v)[32] = g == 0 ? va : vb;
            const int cg = c0 + g * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = ld_shared_f4(b9 + 4u * (cg + j));
              const float4 w01 = ld_shared_f4(w10 + 8u * (cg + j)), w23 = ld_shared_f4(w10 + 8u * (cg + j + 2));
              float a0, a1, a2, a3;
              xg_upk2(xg_add2(xg_pk2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), xg_pk2(bb.x, bb.y)), a0, a1);
              xg_upk2(xg_add2(xg_pk2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), xg_pk2(bb.z, bb.w)), a2, a3);
              a0 = fmaxf(a0, 0.0f); a1 = fmaxf(a1, 0.0f); a2 = fmaxf(a2, 0.0f); a3 = fmaxf(a3, 0.0f);
              l01 = xg_fma2(xg_pk2(a0, a0), xg_pk2(w01.x, w01.y), l01);
              l01 = xg_fma2(xg_pk2(a1, a1), xg_pk2(w01.z, w01.w), l01);
              l01 = xg_fma2(xg_pk2(a2, a2), xg_pk2(w23.x, w23.y), l01);
              l01 = xg_fma2(xg_pk2(a3, a3), xg_pk2(w23.z, w23.w), l01);
            }
          }
        }
        xg_upk2(l01, l0, l1);
        if (half == 1) st_shared_f2(lx + 8u * row, l0, l1);
        named_bar_sync(1 + q, 64);                     // the two P2 warps of this lane quarter
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
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<512>(tmem_base);
}

}  // namespace t3d
