// Split-precision ("f16x2") variant of the fused per-point MLP chain + max-pool (chain_max.cuh): same reference
// functions (sunrgbd_detection/semisup_models.py:76-97, :172-189, :224-245, :354-376; models/model_util.py:300-316),
// same work decomposition and warp roles, but fp32-accurate on the tensor cores:
//
//   every operand is a pair of fp16 images  x = hi + lo  (hi = 11 significant bits of x, lo = fp16(x - hi), 22+ bits
//   together; weights are split once at pack time, activations in the epilogue that produces them), and a layer issues
//   three tcgen05.mma products per K-block into ONE fp32 TMEM accumulator:   hi.lo + lo.hi + hi.hi
//   (lo.lo is below 2^-22 relative and dropped).
//
// Operands are pre-scaled by powers of two so that the lo images stay in the fp16 normal range: activations carry
// kX2ActScale, each layer's weights carry 2^s chosen at pack time (max |W| 2^s in [2^9, 2^10)); the epilogue undoes the
// weight scale with one FFMA (x = acc * inv + bias * kX2ActScale).  `inv` also carries a first-order correction of the
// tensor core's round-toward-zero accumulation (t3d_api.cu: x2_debias; tests/numerics_split_study.py).
// The tensor core truncates at every K=16 accumulation step; per K-block the two small products are issued before the
// main one (chunk stream per (n-block, K-block) = [lo][hi], MMAs = a_hi.w_lo, a_lo.w_hi, a_hi.w_hi), so every chunk leaves
// the ring right after its last product.  (Issuing the small products of TWO K-blocks first -- kX2Group = 2 -- halves the
// truncation steps at partial magnitude but holds the hi chunks in the ring twice as long: measured 97.6 - 98.5 % mask-exact
// either way (profiles/r02_mask_exactness.txt: the de-bias factor removes the mean effect), and the ring stalls cost
// ~5 k cycles per tile, so the streaming order won.)
// One 128-point tile per iteration (the hi + lo images double every activation buffer); for the 256-wide box chains the
// final-layer operand aliases the two hidden buffers.  bf16 counterpart: 62 % tensor-pipe active with one product;
// here the pipe has three products per epilogue.
#pragma once
#include "common.cuh"
#include "chain_max.cuh"

namespace t3d {

constexpr float kX2ActScale = 16.0f;
constexpr int kX2Group = 1;               // K-blocks per small-first group (hi chunks of a group are resident together)
constexpr int kX2Stages = 5;              // 16 KB weight-ring stages (131 KB of activation images leave room for five)

template <typename S> __host__ __device__ constexpr int x2_buf_width(int b) { return S::BUF_BYTES(b) / (256 * S::NSUB); }
template <typename S> __host__ __device__ constexpr int x2_num_chunks() { return 2 * chain_num_chunks<S>(); }
// fp32 tail of the arena: [W1 * As][b1 * As][hidden biases * As][final bias][inv scale of every MMA layer (NH + 1), padded to 4]
template <typename S> __host__ __device__ constexpr int x2_num_floats() {
  return S::CIN * S::C1 + S::C1 + chain_hidden_bias_count<S>() + S::FC + 4;
}
template <typename S> __host__ __device__ constexpr size_t x2_arena_bytes() {
  return (size_t)x2_num_chunks<S>() * kChunkBytes + sizeof(float) * (size_t)x2_num_floats<S>();
}

template <typename S> struct X2Smem {
  static constexpr int W0 = x2_buf_width<S>(0), W1w = x2_buf_width<S>(1), W2 = x2_buf_width<S>(2);
  static constexpr bool ALIAS2 = (W0 + W1w + W2) * 512 > 131072;      // box chains: final operand over the hidden buffers
  static constexpr int BUF0 = 0;
  static constexpr int BUF1 = BUF0 + 512 * W0;
  static constexpr int BUF2 = ALIAS2 ? 0 : BUF1 + 512 * W1w;
  static constexpr int ACT_END = ALIAS2 ? 512 * ((W0 + W1w) > W2 ? (W0 + W1w) : W2) : BUF2 + 512 * W2;
  static constexpr int RING = ACT_END;
  static constexpr int W1 = RING + kX2Stages * kChunkBytes;              // fp32 [CIN][C1]
  static constexpr int B1 = W1 + 4 * S::CIN * S::C1;
  static constexpr int HB = B1 + 4 * S::C1;                                  // hidden biases
  static constexpr int BARS = (HB + 4 * chain_hidden_bias_count<S>() + 15) / 16 * 16;
  // barriers: ring_full[4], ring_empty[4], acc_full[3], acc_empty[3], act_ready[4], front_free
  static constexpr int NBARS = 2 * kX2Stages + 6 + 4 + 1;
  static constexpr int TMEM_SLOT = BARS + 8 * NBARS;
  static constexpr int TOTAL = TMEM_SLOT + 16;
  static constexpr int buf_off(int b) { return b == 0 ? BUF0 : (b == 1 ? BUF1 : BUF2); }
  // the buffer layer-0's input lives in may be overwritten by the front warps once the MMAs of this layer are complete
  static constexpr int FRONT_FREE = ALIAS2 ? S::NH : S::FRONT_FREE_LAYER;   // NH = after the final layer
};

template <int KIND>
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kChainThreads, 1) chain_max_x2_kernel(const ChainArgs args) {
  using S = ChainSpec<KIND>;
  using L = X2Smem<S>;
  constexpr int NMT = S::FC / 128;
  constexpr int NCH = x2_num_chunks<S>();
  static_assert(L::TOTAL + 1024 <= 232448, "chain_max_x2: shared memory budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  constexpr uint16_t kAllCtas = (1u << kClusterSize) - 1;

  const uint32_t bar0 = sbase + L::BARS;
  auto ring_full = [&](int s) { return bar0 + 8u * s; };
  auto ring_empty = [&](int s) { return bar0 + 8u * (kX2Stages + s); };
  auto acc_full = [&](int r) { return bar0 + 8u * (2 * kX2Stages + r); };
  auto acc_empty = [&](int r) { return bar0 + 8u * (2 * kX2Stages + 3 + r); };
  auto act_ready = [&](int a) { return bar0 + 8u * (2 * kX2Stages + 6 + a); };
  const uint32_t front_free = bar0 + 8u * (2 * kX2Stages + 10);
  // TMEM regions: final R0 = cols 0..127, R1 = 128..255 (ping-pong over channel tiles), hidden R2 = 256..511
  auto region_col = [&](int r) -> uint32_t { return (uint32_t)(r * 128); };

  const int tiles_per_frustum = (args.N + 127) / 128;
  const int num_tiles = args.tiles ? *args.num_tiles_ptr : args.B * tiles_per_frustum;
  const int ncl = gridDim.x / kClusterSize, cl = blockIdx.x / kClusterSize;
  const int cbegin = (int)(((long long)num_tiles * cl) / ncl);
  const int cend = (int)(((long long)num_tiles * (cl + 1)) / ncl);
  const int iters = (cend - cbegin + kClusterSize - 1) / kClusterSize;
  auto tile_of = [&](int i) { return min(cbegin + i * kClusterSize + (int)crank, cend - 1); };
  auto tile_info = [&](int t, int& fr, int& start, int& npts) {
    if (args.tiles) { int4 d = args.tiles[t]; fr = d.x; start = d.y; npts = d.z; }
    else { fr = t / tiles_per_frustum; start = (t % tiles_per_frustum) * 128; npts = min(128, args.N - start); }
  };
  const float* ftail = reinterpret_cast<const float*>(args.arena + (size_t)NCH * kChunkBytes);

  // ------------------------------------------------------------------ setup
  if (threadIdx.x == 0) {
    for (int s = 0; s < kX2Stages; ++s) { mbar_init(ring_full(s), 1); mbar_init(ring_empty(s), kClusterSize); }
    for (int r = 0; r < 3; ++r) { mbar_init(acc_full(r), 1); mbar_init(acc_empty(r), 8); }
    for (int a = 0; a < 4; ++a) mbar_init(act_ready(a), a == 0 ? 4 : 8);
    mbar_init(front_free, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(sbase + L::TMEM_SLOT);
  {
    float* fdst = reinterpret_cast<float*>(smem + L::W1);
    constexpr int NF = S::CIN * S::C1 + S::C1 + chain_hidden_bias_count<S>();
    for (int i = threadIdx.x; i < NF; i += blockDim.x) fdst[i] = ftail[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base_v = *reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_SLOT);
  const uint32_t tmem_base = tmem_base_v;

  if (warp == 0) {
    // ================================================================ weight producer (half of every chunk, multicast)
    if (lane == 0) {
      constexpr uint32_t kHalf = kChunkBytes / kClusterSize;
      uint32_t it = 0;
      for (int i = 0; i < iters; ++i) {
        for (int c = 0; c < NCH; ++c, ++it) {
          const int s = it % kX2Stages;
          mbar_wait(ring_empty(s), ((it / kX2Stages) & 1) ^ 1);
          mbar_arrive_expect_tx(ring_full(s), kChunkBytes);
          bulk_g2s_mc(sbase + L::RING + s * kChunkBytes + crank * kHalf, args.arena + (size_t)c * kChunkBytes + crank * kHalf,
                      kHalf, ring_full(s), kAllCtas);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (whole warp converged, one elected lane issues)
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base_v, 0);
    uint32_t it = 0;
    uint32_t acc_cnt[3] = {0, 0, 0};
    Tracer tr; tr.init(lane == 0 ? args.trace : nullptr, 1);
    // one product block: four K=16 steps over a 64-wide K-block.  x_addr / y_addr are the smem operands in (A, B) order.
    auto mma4 = [&](uint32_t d, uint32_t x_addr, uint32_t y_addr, uint32_t idesc, bool first) {
      const uint64_t xd = make_sdesc_k128(x_addr), yd = make_sdesc_k128(y_addr);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_w(d, xd + 2u * k, yd + 2u * k, idesc, (first && k == 0) ? 0u : 1u);
    };
    // all K-blocks of one output block: act = base of the activation buffer (hi image, lo image at +lo_off), kbn K-blocks;
    // swapped = weights are the A operand (final layer).  Consumes 2*kbn ring chunks.
    auto layer_block = [&](uint32_t d, uint32_t act, uint32_t lo_off, int kbn, uint32_t idesc, bool swapped) {
      bool first = true;
      for (int g0 = 0; g0 < kbn; g0 += kX2Group) {
        const int gn = min(kX2Group, kbn - g0);
        for (int j = 0; j < gn; ++j, ++it) {                       // lo chunks: a_hi . w_lo
          const int s = it % kX2Stages;
          mbar_wait_w(ring_full(s), (it / kX2Stages) & 1);
          tc_fence_after();
          const uint32_t w = sbase + L::RING + s * kChunkBytes, a = act + (g0 + j) * 16384;
          if (swapped) mma4(d, w, a, idesc, first); else mma4(d, a, w, idesc, first);
          first = false;
          umma_commit_mc_w(ring_empty(s), kAllCtas);
        }
        for (int j = 0; j < gn; ++j) {                             // hi chunks of the group: a_lo . w_hi
          const int s = (it + j) % kX2Stages;
          mbar_wait_w(ring_full(s), ((it + j) / kX2Stages) & 1);
          tc_fence_after();
          const uint32_t w = sbase + L::RING + s * kChunkBytes, a = act + lo_off + (g0 + j) * 16384;
          if (swapped) mma4(d, w, a, idesc, false); else mma4(d, a, w, idesc, false);
        }
        for (int j = 0; j < gn; ++j, ++it) {                       // main products: a_hi . w_hi
          const int s = it % kX2Stages;
          const uint32_t w = sbase + L::RING + s * kChunkBytes, a = act + (g0 + j) * 16384;
          if (swapped) mma4(d, w, a, idesc, false); else mma4(d, a, w, idesc, false);
          umma_commit_mc_w(ring_empty(s), kAllCtas);
        }
      }
    };
    for (int i = 0; i < iters; ++i) {
      const uint32_t tpar = i & 1;
      int fr, start, npts;
      tile_info(tile_of(i), fr, start, npts);
      tr.mark(0x10);
#pragma unroll
      for (int l = 0; l < S::NH; ++l) {
        const uint32_t a_buf = sbase + L::buf_off(S::ACT_BUF(l));
        const int nbn = (S::HN(l) + 127) / 128, kbn = S::HK(l) / 64;
        mbar_wait_w(act_ready(l), tpar);
        tr.mark(0x20 + l);
        if (S::EMIT_LAYER >= 0 && l == S::EMIT_LAYER + 1 && args.emit != nullptr) {
          // point_feat of this tile: hi image then lo image ([128 x 64] fp16 K-major SW128 each), 32 KB per tile
          const size_t tidx = (size_t)fr * tiles_per_frustum + start / 128;
          if (lane == 0) bulk_s2g(reinterpret_cast<uint8_t*>(args.emit) + tidx * 32768, a_buf, 32768);
          __syncwarp();
        }
        mbar_wait_w(acc_empty(2), (acc_cnt[2] & 1) ^ 1);
        tc_fence_after();
        for (int nb = 0; nb < nbn; ++nb) {
          const int ncols = min(128, S::HN(l) - nb * 128);
          layer_block(tm + region_col(2) + nb * 128, a_buf, (uint32_t)kbn * 16384u, kbn, make_idesc_f16(128, ncols), false);
        }
        umma_commit_w(acc_full(2));
        acc_cnt[2]++;
        tr.mark(0x30 + l);
        if (l == L::FRONT_FREE) {
          if (S::EMIT_LAYER >= 0 && args.emit != nullptr) { if (lane == 0) bulk_wait_read_all(); __syncwarp(); }
          umma_commit_w(front_free);
        }
      }
      // final layer: channels on M, the tile's 128 points on N
      mbar_wait_w(act_ready(S::NH), tpar);
      tr.mark(0x40);
      tc_fence_after();
      const uint32_t b_buf = sbase + L::buf_off(S::ACT_BUF(S::NH));
      for (int mt = 0; mt < NMT; ++mt) {
        const int r = mt & 1;
        mbar_wait_w(acc_empty(r), (acc_cnt[r] & 1) ^ 1);
        tr.mark(0x50 + mt);
        tc_fence_after();
        layer_block(tm + region_col(r), b_buf, (uint32_t)(S::FK / 64) * 16384u, S::FK / 64, make_idesc_f16(128, 128), true);
        umma_commit_w(acc_full(r));
        acc_cnt[r]++;
        tr.mark(0x60 + mt);
      }
      if (L::FRONT_FREE == S::NH) umma_commit_w(front_free);
    }
    if (S::EMIT_LAYER >= 0 && args.emit != nullptr && lane == 0) bulk_wait_all();
  } else if (warp >= 4 && warp < 12) {
    // ================================================================ epilogue warps: TMEM lane quarter = warp&3, column half = (warp-4)>>2
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    uint32_t acc_cnt[3] = {0, 0, 0};
    float run_max[NMT];
#pragma unroll
    for (int i = 0; i < NMT; ++i) run_max[i] = -3.0e38f;
    const uint32_t hbias = sbase + L::HB;
    const float* fbias = ftail + S::CIN * S::C1 + S::C1 + chain_hidden_bias_count<S>();
    const float* scales = fbias + S::FC;
    Tracer tr; tr.init((warp == 4 && lane == 0) ? args.trace : nullptr, 2);

    // 32 accumulator columns [c0, c0+32) of this thread's row: x = acc * inv + bias, ReLU, f16 hi / lo, into the next operand
    auto store_group = [&](const uint32_t (&v)[32], float inv, uint32_t bias, int c0, uint32_t o_buf, uint32_t lo_off) {
      uint32_t ph[16], pl[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = ld_shared_f4(bias + 4u * (c0 + 4 * j));
        split_f16x2_relu(fmaf(__uint_as_float(v[4 * j]), inv, b4.x), fmaf(__uint_as_float(v[4 * j + 1]), inv, b4.y), ph[2 * j], pl[2 * j]);
        split_f16x2_relu(fmaf(__uint_as_float(v[4 * j + 2]), inv, b4.z), fmaf(__uint_as_float(v[4 * j + 3]), inv, b4.w), ph[2 * j + 1],
                         pl[2 * j + 1]);
      }
      const int kb = c0 >> 6, j0 = (c0 & 63) >> 3;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const uint32_t o = o_buf + kb * 16384 + sw128_offset(row, j0 + jj);
        st_shared_v4(o, ph[4 * jj], ph[4 * jj + 1], ph[4 * jj + 2], ph[4 * jj + 3]);
        st_shared_v4(o + lo_off, pl[4 * jj], pl[4 * jj + 1], pl[4 * jj + 2], pl[4 * jj + 3]);
      }
    };

    for (int i = 0; i < iters; ++i) {
      int fr, start, npts;
      tile_info(tile_of(i), fr, start, npts);
      int hb_off = 0;
      tr.mark(0x10);
#pragma unroll
      for (int l = 0; l < S::NH; ++l) {
        const uint32_t o_buf = sbase + L::buf_off(S::ACT_BUF(l + 1));
        const uint32_t lo_off = (uint32_t)(S::HN(l) / 64) * 16384u;
        const int span = S::HN(l) / 2, cbeg = half * span;
        const float inv = scales[l];
        mbar_wait(acc_full(2), acc_cnt[2] & 1);
        acc_cnt[2]++;
        tr.mark(0x20 + l);
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_sel + region_col(2);
#pragma unroll 1
        for (int c0 = cbeg; c0 < cbeg + span; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          store_group(v, inv, hbias + 4u * hb_off, c0, o_buf, lo_off);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(act_ready(l + 1)); mbar_arrive(acc_empty(2)); }
        tr.mark(0x30 + l);
        hb_off += S::HN(l);
      }
      // final layer: this thread owns channel mt*128+row; columns [half*64, +64) are points
#pragma unroll
      for (int mt = 0; mt < NMT; ++mt) {
        const int r = mt & 1;
        mbar_wait(acc_full(r), acc_cnt[r] & 1);
        acc_cnt[r]++;
        tr.mark(0x50 + mt);
        tc_fence_after();
        float m = run_max[mt];
        const uint32_t taddr = tmem_base + lane_sel + region_col(r) + half * 64;
        uint32_t va[32], vb[32];
        tmem_ld32(taddr, va);
        tmem_ld32(taddr + 32, vb);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) m = fmax3(m, __uint_as_float(va[2 * j]), __uint_as_float(va[2 * j + 1]));
#pragma unroll
        for (int j = 0; j < 16; ++j) m = fmax3(m, __uint_as_float(vb[2 * j]), __uint_as_float(vb[2 * j + 1]));
        run_max[mt] = m;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(r));
        tr.mark(0x60 + mt);
      }
      bool flush = (i + 1 == iters);
      if (!flush) { int f2, s2, n2; tile_info(tile_of(i + 1), f2, s2, n2); flush = (f2 != fr); }
      if (flush) {
        const float invf = scales[S::NH];
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt) {
          const int ch = mt * 128 + row;
          const float v = fmaxf(fmaf(run_max[mt], invf, fbias[ch]), 0.0f);
          atomicMax(reinterpret_cast<int*>(args.out + (size_t)fr * S::FC + ch), __float_as_int(v));
          run_max[mt] = -3.0e38f;
        }
      }
    }
  } else if (warp >= 12) {
    // ================================================================ front warps: load points, layer 1 on CUDA cores (fp32), split
    const int p = threadIdx.x - 384;               // 0..127 = row of the tile
    const uint32_t W1 = sbase + L::W1, B1 = sbase + L::B1;
    const uint32_t o_buf = sbase + L::buf_off(S::ACT_BUF(0));
    constexpr uint32_t lo_off = (uint32_t)(S::C1 / 64) * 16384u;
    for (int i = 0; i < iters; ++i) {
      int fr, start, npts;
      tile_info(tile_of(i), fr, start, npts);
      if (i > 0) mbar_wait(front_free, (i - 1) & 1);
      float cx = 0.f, cy = 0.f, cz = 0.f;
      if (args.center) { cx = args.center[fr * 3 + 0]; cy = args.center[fr * 3 + 1]; cz = args.center[fr * 3 + 2]; }
      float bc[3] = {0, 0, 0}, hl = 0, hw = 0, hh = 0, ct = 1, st = 0;
      if (S::BOXPC) {
        bc[0] = args.box_center[fr * 3 + 0]; bc[1] = args.box_center[fr * 3 + 1]; bc[2] = args.box_center[fr * 3 + 2];
        hl = 0.5f * args.box_dims[fr * 3 + 0]; hw = 0.5f * args.box_dims[fr * 3 + 1]; hh = 0.5f * args.box_dims[fr * 3 + 2];
        sincosf(args.box_orient[fr], &st, &ct);
      }
      int j = min(p, npts - 1);                    // padded rows duplicate the last valid point
      j += start;
      const int src = args.idx ? args.idx[(size_t)fr * args.idx_stride + j] : j;
      const float* pp = args.pc + ((size_t)fr * args.N + src) * args.C;
      float x[S::CIN];
#pragma unroll
      for (int k = 0; k < S::CRAW; ++k) x[k] = pp[k];
      x[0] -= cx; x[1] -= cy; x[2] -= cz;
      if (S::BOXPC) {
        const float dx = x[0] - bc[0], dy = x[1] - bc[1], dz = x[2] - bc[2];
        const float xr = ct * dx - st * dz, zr = st * dx + ct * dz;
        x[S::CRAW + 0] = hl - xr; x[S::CRAW + 1] = hl + xr;
        x[S::CRAW + 2] = hh - dy; x[S::CRAW + 3] = hh + dy;
        x[S::CRAW + 4] = hw - zr; x[S::CRAW + 5] = hw + zr;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < S::C1; c0 += 8) {
        float a[8];
        {
          const float4 b0 = ld_shared_f4(B1 + 4u * c0), b1 = ld_shared_f4(B1 + 4u * (c0 + 4));
          a[0] = b0.x; a[1] = b0.y; a[2] = b0.z; a[3] = b0.w; a[4] = b1.x; a[5] = b1.y; a[6] = b1.z; a[7] = b1.w;
        }
#pragma unroll
        for (int k = 0; k < S::CIN; ++k) {
          const float4 w0 = ld_shared_f4(W1 + 4u * (k * S::C1 + c0));
          const float4 w1 = ld_shared_f4(W1 + 4u * (k * S::C1 + c0 + 4));
          a[0] = fmaf(x[k], w0.x, a[0]); a[1] = fmaf(x[k], w0.y, a[1]); a[2] = fmaf(x[k], w0.z, a[2]); a[3] = fmaf(x[k], w0.w, a[3]);
          a[4] = fmaf(x[k], w1.x, a[4]); a[5] = fmaf(x[k], w1.y, a[5]); a[6] = fmaf(x[k], w1.z, a[6]); a[7] = fmaf(x[k], w1.w, a[7]);
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_f16x2_relu(a[2 * e], a[2 * e + 1], h[e], l[e]);
        const uint32_t o = o_buf + (c0 >> 6) * 16384 + sw128_offset(p, (c0 & 63) >> 3);
        st_shared_v4(o, h[0], h[1], h[2], h[3]);
        st_shared_v4(o + lo_off, l[0], l[1], l[2], l[3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(act_ready(0));
    }
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

}  // namespace t3d
