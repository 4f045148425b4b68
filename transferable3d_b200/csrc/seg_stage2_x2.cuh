// Instance-segmentation stage 2 in split precision ("f16x2"; see chain_x2.cuh for the operand format): point_feat(64) ->
// conv6'(512, + per-frustum gbias) -> conv7(256) -> conv8(128) -> conv9(128) -> conv10(2) = mask logits
// (sunrgbd_detection/semisup_models.py:107-135, eval mode, BN folded, conv6's global half folded into gbias).
//
// One 128-point tile at a time per CTA (points on the UMMA M dimension, channels on N).  Every activation passes to the
// next layer through two 32 KB K-block slots (one K-block = 64 channels x 128 points as an fp16 hi image + lo image):
//   conv6' block nb (128 channels, K = 64)  -> e6 -> 2 K-blocks -> conv7 accumulates them (8 K-blocks in all)
//   conv7 (256 channels)                    -> e7 -> 4 K-blocks streamed through the slots -> conv8
//   conv8 (128 channels)                    -> e8 -> 2 K-blocks (both slots)              -> conv9 -> e9 + conv10 (fp32)
// so the 512- and 256-wide activations never exist as a whole.  Three products per K-block (hi.lo, lo.hi, hi.hi) into
// one fp32 accumulator, the two small ones first; the first conv6' block of tile i+1 is issued under the e7 epilogue of
// tile i.  16 epilogue warps (4 lane quarters x 4 column quarters of a K-block).
// TMEM: R6 = cols 0..127 (conv6' block), R89 = 128..255 (conv8, conv9), R7 = 256..511 (conv7).
// The two CTAs of a cluster are one tcgen05 cta_group::2 pair (as in seg_stage2_pair.cuh): M = 256 = their two tiles, every weight
// chunk (fp16 hi / lo image, 16 KB) SPLIT between the two shared memories (rows 64 r .. 64 r + 63 in CTA r: an 8 KB ring stage,
// 10 stages), the leader issues every MMA / commit (multicast to both CTAs' barriers), the peer's epilogue warps arrive remotely on
// the leader's barriers and two relay lanes forward its TMA completions.  The input tile (stage 1's hi + lo point_feat images,
// 32 KB) and gbias are double-buffered.
#pragma once
#include "common.cuh"
#include "chain_max.cuh"
#include "chain_x2.cuh"

namespace t3d {

constexpr int kSeg2XChunks = 52;   // per tile: 8 (W6') + 32 (W7) + 8 (W8) + 4 (W9); consumption order in t3d_pack_seg2_x2
// fp32 tail: [b7*As 256][b8*As 128][b9*As 128][W10/As 128x2][b10 2][inv6 inv7 inv8 inv9][pad 2]
constexpr int kSeg2XFloats = 256 + 128 + 128 + 256 + 2 + 4 + 2;
constexpr size_t kSeg2XArenaBytes = (size_t)kSeg2XChunks * kChunkBytes + sizeof(float) * kSeg2XFloats;

struct Seg2XArgs {
  const uint8_t* point_feat;         // stage-1 emit: per 128-point tile [hi image 16 KB][lo image 16 KB]
  const float* gbias;                // [B, 512] fp32, already multiplied by kX2ActScale
  const uint8_t* arena;
  float* logits;                     // [B, N, 2]
  int B, N;
  unsigned long long* trace;
};

struct Seg2XSmem {
  static constexpr int STAGES = 10;
  static constexpr int STAGE_BYTES = kChunkBytes / 2;   // this CTA's 64 rows of a chunk
  static constexpr int IN = 0;                          // 2 x 32 KB
  static constexpr int SL = 65536;                      // 2 slots x 32 KB
  static constexpr int RING = SL + 65536;
  static constexpr int GB = RING + STAGES * STAGE_BYTES;   // 2 x 512 fp32
  static constexpr int FL = GB + 2 * 2048;
  static constexpr int LX = FL + ((kSeg2XFloats * 4 + 15) / 16) * 16;   // [128][3][2] fp32 partial logits
  static constexpr int BARS = LX + 128 * 24;
  // ring_full[S] ring_empty[S] in_ready[2] in_free[2] r6_full r6_empty sl_ready[2] sl_free[2] r7_full r7_empty r89_full r89_empty
  static constexpr int NBARS = 2 * STAGES + 14;
  static constexpr int TMEM_SLOT = BARS + 8 * NBARS;
  static constexpr int TOTAL = TMEM_SLOT + 16;
};
static_assert(Seg2XSmem::TOTAL + 1024 <= 232448, "seg_stage2_x2: shared memory budget");

constexpr int kSeg2XThreads = 640;   // warp 0 weight producer, 1 MMA (leader) / ring relay (peer), 2 TMEM alloc (+ input relay in the peer),
                                     // 3 input producer, 4-19 epilogue
static_assert(kClusterSize == 2, "seg_stage2_x2: one CTA pair per cluster");

__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kSeg2XThreads, 1) seg_stage2_x2_kernel(const Seg2XArgs args) {
  using L = Seg2XSmem;
  constexpr int ST = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  constexpr uint16_t kBoth = 3;
  constexpr uint32_t kStageBytes = L::STAGE_BYTES;

  const uint32_t bar0 = sbase + L::BARS;
  auto ring_full = [&](int s) { return bar0 + 8u * s; };
  auto ring_empty = [&](int s) { return bar0 + 8u * (ST + s); };
  constexpr int B0 = 2 * ST;
  auto in_ready = [&](int b) { return bar0 + 8u * (B0 + 0 + b); };
  auto in_free = [&](int b) { return bar0 + 8u * (B0 + 2 + b); };
  const uint32_t r6_full = bar0 + 8u * (B0 + 4), r6_empty = bar0 + 8u * (B0 + 5);
  auto sl_ready = [&](int b) { return bar0 + 8u * (B0 + 6 + b); };
  auto sl_free = [&](int b) { return bar0 + 8u * (B0 + 8 + b); };
  const uint32_t r7_full = bar0 + 8u * (B0 + 10), r7_empty = bar0 + 8u * (B0 + 11);
  const uint32_t r89_full = bar0 + 8u * (B0 + 12), r89_empty = bar0 + 8u * (B0 + 13);
  constexpr uint32_t kR6 = 0, kR89 = 128, kR7 = 256;

  const int tiles_per_frustum = (args.N + 127) / 128;
  const int num_tiles = args.B * tiles_per_frustum;
  const int ncl = gridDim.x / kClusterSize, cl = blockIdx.x / kClusterSize;
  const int cbegin = (int)(((long long)num_tiles * cl) / ncl);
  const int cend = (int)(((long long)num_tiles * (cl + 1)) / ncl);
  const int iters = (cend - cbegin + kClusterSize - 1) / kClusterSize;
  auto tile_of = [&](int i) { return min(cbegin + i * kClusterSize + (int)crank, cend - 1); };
  // epilogue -> MMA-warp signals: the MMA warp lives in the leader CTA
  auto arrive_mma = [&](uint32_t bar) { if (leader) mbar_arrive(bar); else mbar_arrive_remote(bar, 0); };

  if (threadIdx.x == 0) {
    // the leader's ring_full / in_ready also count the peer's forwarded completion; its epilogue-side barriers count the 16
    // warps of both CTAs; everything the MMA warp signals is one multicast commit per CTA
    const uint32_t fwd = leader ? 2 : 1;
    for (int s = 0; s < ST; ++s) { mbar_init(ring_full(s), fwd); mbar_init(ring_empty(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(in_ready(b), fwd); mbar_init(in_free(b), 1); mbar_init(sl_ready(b), 32); mbar_init(sl_free(b), 1); }
    mbar_init(r6_full, 1); mbar_init(r6_empty, 32);
    mbar_init(r7_full, 1); mbar_init(r7_empty, 32);
    mbar_init(r89_full, 1); mbar_init(r89_empty, 32);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<512>(sbase + L::TMEM_SLOT);
  {
    const float* fsrc = reinterpret_cast<const float*>(args.arena + (size_t)kSeg2XChunks * kChunkBytes);
    float* fdst = reinterpret_cast<float*>(smem + L::FL);
    for (int i = threadIdx.x; i < kSeg2XFloats; i += blockDim.x) fdst[i] = fsrc[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_SLOT);

  if (warp == 0) {
    // ================================================================ weight producer: this CTA's 64 rows of every chunk, kept local
    if (lane == 0) {
      uint32_t it = 0;
      auto push = [&](int c) {
        const int s = it % ST;
        mbar_wait(ring_empty(s), ((it / ST) & 1) ^ 1);
        mbar_arrive_expect_tx(ring_full(s), kStageBytes);
        bulk_g2s(sbase + L::RING + s * kStageBytes, args.arena + (size_t)c * kChunkBytes + crank * kStageBytes, kStageBytes, ring_full(s));
        ++it;
      };
      // The stream follows the MMA warp's CONSUMPTION order, not the arena order: the first conv6' block of tile i + 1
      // (arena chunks 0, 1) is issued under the e7 epilogue of tile i, i.e. between conv7 (chunks 2 .. kC8 - 1) and conv8.
      constexpr int kC8 = kSeg2XChunks - 12;          // conv8 (4 K-blocks x [lo, hi]) + conv9 (2 x [lo, hi]) close the arena
      for (int i = 0; i < iters; ++i) {
        if (i == 0) { push(0); push(1); }
        for (int c = 2; c < kC8; ++c) push(c);
        if (i + 1 < iters) { push(0); push(1); }
        for (int c = kC8; c < kSeg2XChunks; ++c) push(c);
      }
    }
  } else if (!leader && warp == 1) {
    // ================================================================ peer: relay the ring's TMA completions to the leader's MMA warp
    if (lane == 0) {
      const uint32_t total = (uint32_t)iters * (uint32_t)kSeg2XChunks;
      for (uint32_t it = 0; it < total; ++it) {
        const int s = it % ST;
        mbar_wait(ring_full(s), (it / ST) & 1);
        mbar_arrive_remote(ring_full(s), 0);
      }
    }
  } else if (!leader && warp == 2) {
    // ================================================================ peer: relay "input tile landed" to the leader's MMA warp
    if (lane == 0) {
      for (int i = 0; i < iters; ++i) {
        mbar_wait(in_ready(i & 1), (i >> 1) & 1);
        mbar_arrive_remote(in_ready(i & 1), 0);
      }
    }
  } else if (warp == 3) {
    // ================================================================ input producer: point_feat tile (hi + lo) + gbias
    if (lane == 0) {
      for (int i = 0; i < iters; ++i) {
        const int t = tile_of(i);
        const int fr = t / tiles_per_frustum;
        const int b = i & 1;
        if (i >= 2) mbar_wait(in_free(b), ((i >> 1) - 1) & 1);
        mbar_arrive_expect_tx(in_ready(b), 32768 + 2048);
        bulk_g2s(sbase + L::IN + b * 32768, args.point_feat + (size_t)t * 32768, 32768, in_ready(b));
        bulk_g2s(sbase + L::GB + b * 2048, args.gbias + (size_t)fr * 512, 2048, in_ready(b));
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader: one instruction drives both SMs; whole warp
    // converged, one elected lane issues)
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t it = 0, n_r6 = 0, n_r89 = 0, n_sl[2] = {0, 0};
    const uint32_t idesc = make_idesc_f16(256, 128);
    Tracer tr; tr.init(lane == 0 ? args.trace : nullptr, 1);
    auto ring_wait = [&](uint32_t j) -> uint32_t {          // chunk j of the stream has landed; returns its smem address
      const uint32_t s = j % ST;
      mbar_wait_w(ring_full(s), (j / ST) & 1);
      return sbase + L::RING + s * kStageBytes;
    };
    auto ring_release = [&](uint32_t j) { umma_commit_pair_w(ring_empty(j % ST), kBoth); };
    auto mma4 = [&](uint32_t d, uint32_t a_addr, uint32_t b_addr, bool first) {
      const uint64_t ad = make_sdesc_k128(a_addr), bd = make_sdesc_k128(b_addr);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_pair_w(d, ad + 2u * k, bd + 2u * k, idesc, (first && k == 0) ? 0u : 1u);
    };
    // one K-block (hi at a, lo at a + 16 KB) against the next [lo, hi] chunk pair: a_hi.w_lo, a_lo.w_hi, a_hi.w_hi
    auto kblock = [&](uint32_t d, uint32_t a, bool first) {
      const uint32_t wl = ring_wait(it);
      tc_fence_after();
      mma4(d, a, wl, first);
      ring_release(it);
      const uint32_t wh = ring_wait(it + 1);
      tc_fence_after();
      mma4(d, a + 16384, wh, false);
      mma4(d, a, wh, false);
      ring_release(it + 1);
      it += 2;
    };
    auto slot_wait = [&](int b) { mbar_wait_w(sl_ready(b), n_sl[b] & 1); n_sl[b]++; tc_fence_after(); };
    bool pre = false;              // conv6' block 0 of the current tile was already issued under the previous tile's tail
    for (int i = 0; i < iters; ++i) {
      const int ib = i & 1;
      const uint32_t in = sbase + L::IN + ib * 32768;
      tr.mark(0x10);
      auto c6 = [&](int nb, uint32_t in_addr) {
        mbar_wait_w(r6_empty, (n_r6 & 1) ^ 1);
        tc_fence_after();
        tr.mark(0x20 + nb);
        kblock(tm + kR6, in_addr, true);
        umma_commit_pair_w(r6_full, kBoth); n_r6++;
      };
      // conv7 K-block kbg (slot kbg & 1): chunks [lo rows 0-127][lo rows 128-255][hi rows 0-127][hi rows 128-255]
      auto c7 = [&](int kbg) {
        const int b = kbg & 1;
        slot_wait(b);
        if (kbg == 0) { mbar_wait_w(r7_empty, (i & 1) ^ 1); tc_fence_after(); }
        if (kbg == 7) umma_commit_pair_w(in_free(ib), kBoth);        // e6(3) has read its gbias, every conv6' MMA of the tile is issued
        tr.mark(0x30 + kbg);
        const uint32_t a = sbase + L::SL + b * 32768;
        const uint32_t wl0 = ring_wait(it), wl1 = ring_wait(it + 1);
        tc_fence_after();
        mma4(tm + kR7, a, wl0, kbg == 0);
        mma4(tm + kR7 + 128, a, wl1, kbg == 0);
        ring_release(it); ring_release(it + 1);
        const uint32_t wh0 = ring_wait(it + 2), wh1 = ring_wait(it + 3);
        tc_fence_after();
        mma4(tm + kR7, a + 16384, wh0, false);
        mma4(tm + kR7, a, wh0, false);
        ring_release(it + 2);
        mma4(tm + kR7 + 128, a + 16384, wh1, false);
        mma4(tm + kR7 + 128, a, wh1, false);
        ring_release(it + 3);
        it += 4;
        umma_commit_pair_w(sl_free(b), kBoth);
        if (kbg == 7) umma_commit_pair_w(r7_full, kBoth);
      };
      if (!pre) {
        mbar_wait_w(in_ready(ib), (i >> 1) & 1);
        tc_fence_after();
        c6(0, in);
      }
      c6(1, in); c7(0); c7(1); c6(2, in); c7(2); c7(3); c6(3, in); c7(4); c7(5); c7(6); c7(7);
      // the tensor pipe would idle while the epilogue works through e7: start the next tile's first conv6' block now
      pre = false;
      if (i + 1 < iters) {
        mbar_wait_w(in_ready((i + 1) & 1), ((i + 1) >> 1) & 1);
        tc_fence_after();
        c6(0, sbase + L::IN + ((i + 1) & 1) * 32768);
        pre = true;
      }
      // conv8: the four K-blocks of the conv7 activation stream through the slots
      for (int kb = 0; kb < 4; ++kb) {
        const int b = kb & 1;
        slot_wait(b);
        if (kb == 0) { mbar_wait_w(r89_empty, (n_r89 & 1) ^ 1); tc_fence_after(); }
        tr.mark(0x40 + kb);
        kblock(tm + kR89, sbase + L::SL + b * 32768, kb == 0);
        umma_commit_pair_w(sl_free(b), kBoth);
      }
      umma_commit_pair_w(r89_full, kBoth); n_r89++;
      // conv9: the two K-blocks of the conv8 activation sit in the two slots
      slot_wait(0);
      mbar_wait_w(r89_empty, (n_r89 & 1) ^ 1);
      tc_fence_after();
      tr.mark(0x50);
      kblock(tm + kR89, sbase + L::SL, true);
      umma_commit_pair_w(sl_free(0), kBoth);
      slot_wait(1);
      kblock(tm + kR89, sbase + L::SL + 32768, false);
      umma_commit_pair_w(r89_full, kBoth); n_r89++;
      umma_commit_pair_w(sl_free(1), kBoth);
      tr.mark(0x51);
    }
  } else if (warp >= 4) {
    // ================================================================ epilogue warps: lane quarter q, column quarter sub of a K-block
    const int q = warp & 3;
    const int sub = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const uint32_t fl = sbase + L::FL;
    const uint32_t b7 = fl, b8 = fl + 4 * 256, b9 = fl + 4 * 384, w10 = fl + 4 * 512, b10 = fl + 4 * 768;
    const float* sc = reinterpret_cast<const float*>(smem + L::FL) + 770;
    const float inv6 = sc[0], inv7 = sc[1], inv8 = sc[2], inv9 = sc[3];
    const uint32_t lx = sbase + L::LX;
    uint32_t n_sl[2] = {0, 0}, n_r6 = 0, n_r89 = 0;
    Tracer tr; tr.init((warp == 4 && lane == 0) ? args.trace : nullptr, 2);

    auto release_acc = [&](uint32_t empty_bar) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_mma(empty_bar);
    };
    // this warp's 16 columns [sub*16, +16) of a K-block: x = acc * inv + bias, ReLU, f16 hi / lo images of slot b
    auto store16 = [&](const uint32_t (&v)[16], float inv, uint32_t bias, int b) {
      mbar_wait(sl_free(b), (n_sl[b] & 1) ^ 1);
      n_sl[b]++;
      uint32_t ph[8], pl[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b4 = ld_shared_f4(bias + 16u * j);
        split_f16x2_relu(fmaf(__uint_as_float(v[4 * j]), inv, b4.x), fmaf(__uint_as_float(v[4 * j + 1]), inv, b4.y), ph[2 * j], pl[2 * j]);
        split_f16x2_relu(fmaf(__uint_as_float(v[4 * j + 2]), inv, b4.z), fmaf(__uint_as_float(v[4 * j + 3]), inv, b4.w), ph[2 * j + 1],
                         pl[2 * j + 1]);
      }
      const uint32_t slot = sbase + L::SL + b * 32768;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const uint32_t o = slot + sw128_offset(row, sub * 2 + jj);
        st_shared_v4(o, ph[4 * jj], ph[4 * jj + 1], ph[4 * jj + 2], ph[4 * jj + 3]);
        st_shared_v4(o + 16384, pl[4 * jj], pl[4 * jj + 1], pl[4 * jj + 2], pl[4 * jj + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) arrive_mma(sl_ready(b));
    };

    for (int i = 0; i < iters; ++i) {
      const int t = tile_of(i);
      const int fr = t / tiles_per_frustum;
      const int start = (t % tiles_per_frustum) * 128;
      const int npts = min(128, args.N - start);
      const uint32_t gb = sbase + L::GB + (i & 1) * 2048;
      mbar_wait(in_ready(i & 1), (i >> 1) & 1);          // gbias of this tile has landed
      tr.mark(0x10);
      // ---- e6: four conv6' blocks, two K-blocks each
      for (int nb = 0; nb < 4; ++nb) {
        uint32_t v0[16], v1[16];
        mbar_wait(r6_full, n_r6 & 1); n_r6++;
        tc_fence_after();
        tr.mark(0x20 + nb);
        tmem_ld16(tmem_base + lane_sel + kR6 + sub * 16, v0);
        tmem_ld16(tmem_base + lane_sel + kR6 + 64 + sub * 16, v1);
        tmem_ld_wait();
        release_acc(r6_empty);
        store16(v0, inv6, gb + 4u * (nb * 128 + sub * 16), 0);
        store16(v1, inv6, gb + 4u * (nb * 128 + 64 + sub * 16), 1);
        tr.mark(0x28 + nb);
      }
      // ---- e7: conv7 activation, K-block by K-block
      mbar_wait(r7_full, i & 1);
      tc_fence_after();
      tr.mark(0x30);
      for (int kb = 0; kb < 4; ++kb) {
        uint32_t v[16];
        tmem_ld16(tmem_base + lane_sel + kR7 + kb * 64 + sub * 16, v);
        tmem_ld_wait();
        if (kb == 3) release_acc(r7_empty);
        store16(v, inv7, b7 + 4u * (kb * 64 + sub * 16), kb & 1);
      }
      tr.mark(0x31);
      // ---- e8
      mbar_wait(r89_full, n_r89 & 1); n_r89++;
      tc_fence_after();
      tr.mark(0x40);
      {
        uint32_t v0[16], v1[16];
        tmem_ld16(tmem_base + lane_sel + kR89 + sub * 16, v0);
        tmem_ld16(tmem_base + lane_sel + kR89 + 64 + sub * 16, v1);
        tmem_ld_wait();
        release_acc(r89_empty);
        store16(v0, inv8, b8 + 4u * (sub * 16), 0);
        store16(v1, inv8, b8 + 4u * (64 + sub * 16), 1);
      }
      tr.mark(0x41);
      // ---- e9 + conv10 (128 -> 2) in fp32: this warp reduces channels [sub*32, +32)
      mbar_wait(r89_full, n_r89 & 1); n_r89++;
      tc_fence_after();
      tr.mark(0x50);
      float l0 = 0.f, l1 = 0.f;
      {
        uint32_t v[32];
        const int c0 = sub * 32;
        tmem_ld32(tmem_base + lane_sel + kR89 + c0, v);
        tmem_ld_wait();
        release_acc(r89_empty);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = ld_shared_f4(b9 + 4u * (c0 + j));
          const float4 w01 = ld_shared_f4(w10 + 8u * (c0 + j)), w23 = ld_shared_f4(w10 + 8u * (c0 + j + 2));
          const float a0 = fmaxf(fmaf(__uint_as_float(v[j]), inv9, bb.x), 0.0f), a1 = fmaxf(fmaf(__uint_as_float(v[j + 1]), inv9, bb.y), 0.0f);
          const float a2 = fmaxf(fmaf(__uint_as_float(v[j + 2]), inv9, bb.z), 0.0f), a3 = fmaxf(fmaf(__uint_as_float(v[j + 3]), inv9, bb.w), 0.0f);
          l0 = fmaf(a0, w01.x, l0); l1 = fmaf(a0, w01.y, l1);
          l0 = fmaf(a1, w01.z, l0); l1 = fmaf(a1, w01.w, l1);
          l0 = fmaf(a2, w23.x, l0); l1 = fmaf(a2, w23.y, l1);
          l0 = fmaf(a3, w23.z, l0); l1 = fmaf(a3, w23.w, l1);
        }
      }
      if (sub != 0) st_shared_f2(lx + 24u * row + 8u * (sub - 1), l0, l1);
      named_bar_sync(1 + q, 128);                    // the four warps of this lane quarter
      if (sub == 0) {
        const float2 o1 = ld_shared_f2(lx + 24u * row), o2 = ld_shared_f2(lx + 24u * row + 8), o3 = ld_shared_f2(lx + 24u * row + 16);
        const float2 bz = ld_shared_f2(b10);
        if (row < npts)
          *reinterpret_cast<float2*>(args.logits + ((size_t)fr * args.N + start + row) * 2) =
              make_float2(((l0 + o1.x) + (o2.x + o3.x)) + bz.x, ((l1 + o1.y) + (o2.y + o3.y)) + bz.y);
      }
      named_bar_sync(1 + q, 128);                    // lx may be overwritten by the next tile only after it was read
      tr.mark(0x51);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<512>(tmem_base);
}

}  // namespace t3d
