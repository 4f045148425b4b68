// Fused per-point MLP chain + max-pool over points on tcgen05/TMEM (sm_100a).
//
// Replaces the conv2d(1x1)+BN+ReLU ... max_pool2d stacks of the reference
// (sunrgbd_detection/semisup_models.py:76-97 inst_seg conv1-5, :172-189 tnet, :224-245 box_est,
// :354-376 box_pc_mask_model; models/model_util.py:300-316 F-PointNet T-Net) in eval mode with BN
// folded into (W,b).  The B x N x C_last activation never reaches HBM.
//
// Work unit: a tile of TILE_PTS = 128*NSUB points of one frustum.
//   layer 1 (Cin=3/6/12 -> C1)       CUDA cores (front warps), bf16 result written to smem in the
//                                     K-major SWIZZLE_128B layout the tensor core reads as operand
//   hidden layers (K -> N<=256)      tcgen05.mma M=128 (points) x N (channels); epilogue warps read
//                                     TMEM, add bias, ReLU, convert to bf16 and write the next operand
//   final layer (K -> CF) + max      swapped operands: M=128 (channels) x N=TILE_PTS (points), so each
//                                     epilogue thread owns one channel and the max over points is a
//                                     register reduction over its TMEM columns
// Weights stream from L2 through a ring of 16 KB stages with cp.async.bulk (TMA engine) in
// pre-swizzled chunk images [<=128 rows x 64 K]; the two CTAs of a cluster each fetch half of every
// chunk and multicast it to both, halving L2 traffic.  mbarriers order producer / MMA / epilogue.
// The two 128-point sub-tiles of a tile are independent chains through the hidden layers and are
// ping-ponged (MMA of one overlaps the epilogue of the other); 8 epilogue warps split the columns.
// Masked stacks run on compacted (masked-in) points only, which is exactly the reference's
// max(act*mask): post-ReLU activations are >= 0 and the output is zero-initialised.
#pragma once
#include "common.cuh"

namespace t3d {

constexpr int kChunkBytes = 16384;        // [128 rows x 64 bf16]
constexpr int kRingStages = 4;

enum ChainKind { CHAIN_SEG1 = 0, CHAIN_TNET = 1, CHAIN_BOX = 2, CHAIN_BOXPC = 3, CHAIN_BOXPCB = 4 };

template <int KIND> struct ChainSpec;
template <> struct ChainSpec<CHAIN_SEG1> {
  static constexpr int CIN = 6, CRAW = 6, C1 = 64, NH = 3, NSUB = 2;
  static constexpr int FK = 128, FC = 1024;
  static constexpr int HK(int i) { return i == 0 ? 64 : i == 1 ? 64 : 64; }
  static constexpr int HN(int i) { return i == 0 ? 64 : i == 1 ? 64 : 128; }
  static constexpr int ACT_BUF(int i) { return i == 0 ? 0 : i == 1 ? 1 : i == 2 ? 0 : 2; }
  static constexpr int BUF_BYTES(int i) { return i == 0 ? 32768 : i == 1 ? 32768 : 65536; }
  static constexpr int FRONT_FREE_LAYER = 2, EMIT_LAYER = 1;
  static constexpr bool BOXPC = false;
};
template <> struct ChainSpec<CHAIN_TNET> {
  static constexpr int CIN = 3, CRAW = 3, C1 = 128, NH = 1, NSUB = 2;
  static constexpr int FK = 128, FC = 256;
  static constexpr int HK(int i) { return i == 0 ? 128 : i == 1 ? 0 : 0; }
  static constexpr int HN(int i) { return i == 0 ? 128 : i == 1 ? 0 : 0; }
  static constexpr int ACT_BUF(int i) { return i == 0 ? 0 : i == 1 ? 1 : i == 2 ? 0 : 0; }
  static constexpr int BUF_BYTES(int i) { return i == 0 ? 65536 : i == 1 ? 65536 : 0; }
  static constexpr int FRONT_FREE_LAYER = 0, EMIT_LAYER = -1;
  static constexpr bool BOXPC = false;
};
template <> struct ChainSpec<CHAIN_BOX> {
  static constexpr int CIN = 3, CRAW = 3, C1 = 128, NH = 2, NSUB = 1;
  static constexpr int FK = 256, FC = 512;
  static constexpr int HK(int i) { return i == 0 ? 128 : i == 1 ? 128 : 0; }
  static constexpr int HN(int i) { return i == 0 ? 128 : i == 1 ? 256 : 0; }
  static constexpr int ACT_BUF(int i) { return i == 0 ? 0 : i == 1 ? 1 : i == 2 ? 2 : 0; }
  static constexpr int BUF_BYTES(int i) { return i == 0 ? 32768 : i == 1 ? 32768 : 65536; }
  static constexpr int FRONT_FREE_LAYER = 0, EMIT_LAYER = -1;
  static constexpr bool BOXPC = false;
};
template <> struct ChainSpec<CHAIN_BOXPC> {
  static constexpr int CIN = 12, CRAW = 6, C1 = 128, NH = 2, NSUB = 1;
  static constexpr int FK = 256, FC = 512;
  static constexpr int HK(int i) { return i == 0 ? 128 : i == 1 ? 128 : 0; }
  static constexpr int HN(int i) { return i == 0 ? 128 : i == 1 ? 256 : 0; }
  static constexpr int ACT_BUF(int i) { return i == 0 ? 0 : i == 1 ? 1 : i == 2 ? 2 : 0; }
  static constexpr int BUF_BYTES(int i) { return i == 0 ? 32768 : i == 1 ? 32768 : 65536; }
  static constexpr int FRONT_FREE_LAYER = 0, EMIT_LAYER = -1;
  static constexpr bool BOXPC = true;
};

// BoxPC representation B (semisup_models.py:400-471): the conv stack sees the raw 6-channel points, no box prologue
template <> struct ChainSpec<CHAIN_BOXPCB> {
  static constexpr int CIN = 6, CRAW = 6, C1 = 128, NH = 2, NSUB = 1;
  static constexpr int FK = 256, FC = 512;
  static constexpr int HK(int i) { return i == 0 ? 128 : i == 1 ? 128 : 0; }
  static constexpr int HN(int i) { return i == 0 ? 128 : i == 1 ? 256 : 0; }
  static constexpr int ACT_BUF(int i) { return i == 0 ? 0 : i == 1 ? 1 : i == 2 ? 2 : 0; }
  static constexpr int BUF_BYTES(int i) { return i == 0 ? 32768 : i == 1 ? 32768 : 65536; }
  static constexpr int FRONT_FREE_LAYER = 0, EMIT_LAYER = -1;
  static constexpr bool BOXPC = false;
};

template <typename S> __host__ __device__ constexpr int chain_num_chunks() {
  int n = 0;
  for (int l = 0; l < S::NH; ++l) n += ((S::HN(l) + 127) / 128) * (S::HK(l) / 64);
  n += (S::FC / 128) * (S::FK / 64);
  return n;
}
template <typename S> __host__ __device__ constexpr int chain_hidden_bias_count() {
  int n = 0;
  for (int l = 0; l < S::NH; ++l) n += S::HN(l);
  return n;
}
// arena = [chunk images][W1 fp32 CIN*C1][b1 C1][hidden biases][final bias FC]
template <typename S> __host__ __device__ constexpr size_t chain_arena_bytes() {
  return (size_t)chain_num_chunks<S>() * kChunkBytes +
         sizeof(float) * (size_t)(S::CIN * S::C1 + S::C1 + chain_hidden_bias_count<S>() + S::FC);
}

struct ChainArgs {
  const float* pc;          // [B, N, C] fp32 points
  int B, N, C;
  const float* center;      // [B,3] subtracted from xyz (null: none)
  const int* idx;           // [B, idx_stride] point indices (null: dense 0..N-1)
  int idx_stride;
  const int* count;         // [B] valid entries of idx per frustum (null: idx_stride, or N when dense)
  const int4* tiles;        // tile table {frustum, start, npts, 0} (null: dense arithmetic)
  const int* num_tiles_ptr; // device count of entries in `tiles`
  const float* box_center;  // BoxPC: [B,3], [B,3] (l,w,h), [B]
  const float* box_dims;
  const float* box_orient;
  const uint8_t* arena;     // packed weights
  float* out;               // [B, FC] fp32, zero-initialised by the caller; atomic max target
  __nv_bfloat16* emit;      // [B*N, HN[EMIT_LAYER]] bf16 (seg1: point_feat) or null
  unsigned long long* trace; // debug timeline of CTA 0 (null: off)
  const uint8_t* rgb;       // wire format of a 6-channel input (null: pc holds all C channels): pc = xyz [B, N, 3] fp32 and
                            // channels 3..5 = rgb[B, N, 3] / 255 (IEEE division: bit-identical to t3d_assemble_points)
};

// smem carve-up (offsets from the 1024-aligned base)
template <typename S> struct ChainSmem {
  static constexpr int TILE = 128 * S::NSUB;
  static constexpr int BUF0 = 0;
  static constexpr int BUF1 = BUF0 + S::BUF_BYTES(0);
  static constexpr int BUF2 = BUF1 + S::BUF_BYTES(1);
  static constexpr int RING = BUF2 + S::BUF_BYTES(2);
  static constexpr int W1 = RING + kRingStages * kChunkBytes;              // fp32 [CIN][C1]
  static constexpr int B1 = W1 + 4 * S::CIN * S::C1;
  static constexpr int HB = B1 + 4 * S::C1;                                  // hidden biases
  static constexpr int BARS = (HB + 4 * chain_hidden_bias_count<S>() + 15) / 16 * 16;
  // barriers: ring_full[4], ring_empty[4], acc_full[3], acc_empty[3], act_ready[4][2], front_free
  static constexpr int NBARS = 2 * kRingStages + 6 + 8 + 1;
  static constexpr int TMEM_SLOT = BARS + 8 * NBARS;
  static constexpr int TOTAL = TMEM_SLOT + 16;
  static constexpr int buf_off(int b) { return b == 0 ? BUF0 : (b == 1 ? BUF1 : BUF2); }
};

constexpr int kChainThreads = 512;   // warp 0 producer, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue, 12-15 front
constexpr int kClusterSize = 2;

template <int KIND>
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kChainThreads, 1) chain_max_kernel(const ChainArgs args) {
  using S = ChainSpec<KIND>;
  using L = ChainSmem<S>;
  constexpr int TILE = L::TILE;
  constexpr int NSUB = S::NSUB;
  constexpr int NMT = S::FC / 128;

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
  auto acc_full = [&](int r) { return bar0 + 8u * (2 * kRingStages + r); };
  auto acc_empty = [&](int r) { return bar0 + 8u * (2 * kRingStages + 3 + r); };
  auto act_ready = [&](int a, int sub) { return bar0 + 8u * (2 * kRingStages + 6 + a * 2 + sub); };
  const uint32_t front_free = bar0 + 8u * (2 * kRingStages + 14);

  // TMEM regions (column bases): NSUB==2: R0=0, R1=256 shared by hidden(sub) and final(mt&1);
  //                              NSUB==1: final R0=0, R1=128, hidden R2=256.
  auto region_col = [&](int r) -> uint32_t { return NSUB == 2 ? (uint32_t)(r * 256) : (r == 2 ? 256u : (uint32_t)(r * 128)); };
  auto hidden_region = [&](int sub) { return NSUB == 2 ? sub : 2; };

  // Tiles: every cluster owns a contiguous range; its two CTAs take alternate tiles and run the SAME
  // number of iterations (the weight ring is shared in lock-step); a CTA without a tile of its own
  // repeats the cluster's last tile (all outputs are idempotent: max / identical stores).
  const int tiles_per_frustum = (args.N + TILE - 1) / TILE;
  const int num_tiles = args.tiles ? *args.num_tiles_ptr : args.B * tiles_per_frustum;
  const int ncl = gridDim.x / kClusterSize, cl = blockIdx.x / kClusterSize;
  const int cbegin = (int)(((long long)num_tiles * cl) / ncl);
  const int cend = (int)(((long long)num_tiles * (cl + 1)) / ncl);
  const int iters = (cend - cbegin + kClusterSize - 1) / kClusterSize;
  auto tile_of = [&](int i) { return min(cbegin + i * kClusterSize + (int)crank, cend - 1); };
  auto tile_info = [&](int t, int& fr, int& start, int& npts) {
    if (args.tiles) { int4 d = args.tiles[t]; fr = d.x; start = d.y; npts = d.z; }
    else { fr = t / tiles_per_frustum; start = (t % tiles_per_frustum) * TILE; npts = min(TILE, args.N - start); }
  };

  // ------------------------------------------------------------------ setup
  if (threadIdx.x == 0) {
    for (int s = 0; s < kRingStages; ++s) { mbar_init(ring_full(s), 1); mbar_init(ring_empty(s), kClusterSize); }
    for (int r = 0; r < 3; ++r) { mbar_init(acc_full(r), 1); mbar_init(acc_empty(r), 8); }
    for (int a = 0; a < 4; ++a)
      for (int sub = 0; sub < 2; ++sub) mbar_init(act_ready(a, sub), a == 0 ? 4 : 8);
    mbar_init(front_free, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(sbase + L::TMEM_SLOT);
  {   // layer-1 weights + hidden biases -> smem (fp32)
    const float* fsrc = reinterpret_cast<const float*>(args.arena + (size_t)chain_num_chunks<S>() * kChunkBytes);
    float* fdst = reinterpret_cast<float*>(smem + L::W1);
    constexpr int NF = S::CIN * S::C1 + S::C1 + chain_hidden_bias_count<S>();
    for (int i = threadIdx.x; i < NF; i += blockDim.x) fdst[i] = fsrc[i];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // peer barriers are initialised before any multicast reaches them
  tc_fence_after();
  const uint32_t tmem_base_v = *reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_SLOT);
  const uint32_t tmem_base = tmem_base_v;

  if (warp == 0) {
    // ================================================================ weight producer (half of every chunk, multicast)
    if (lane == 0) {
      constexpr int NCH = chain_num_chunks<S>();
      constexpr uint32_t kHalf = kChunkBytes / kClusterSize;
      uint32_t it = 0;
      Tracer tr; tr.init(args.trace, 0);
      for (int i = 0; i < iters; ++i) {
        for (int c = 0; c < NCH; ++c, ++it) {
          const int s = it % kRingStages;
          mbar_wait(ring_empty(s), ((it / kRingStages) & 1) ^ 1);
          tr.mark(1);
          mbar_arrive_expect_tx(ring_full(s), kChunkBytes);
          bulk_g2s_mc(sbase + L::RING + s * kChunkBytes + crank * kHalf, args.arena + (size_t)c * kChunkBytes + crank * kHalf,
                      kHalf, ring_full(s), kAllCtas);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer: the whole warp runs this loop converged,
    // one elected lane issues (umma_*_w, common.cuh); every value below is warp-uniform
    {
      const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_v, 0);
      uint32_t it = 0;                       // ring iteration of the first chunk of the current layer
      uint32_t acc_cnt[3] = {0, 0, 0};       // uses of each TMEM region so far
      Tracer tr; tr.init(lane == 0 ? args.trace : nullptr, 1);
      for (int i = 0; i < iters; ++i) {
        const uint32_t tpar = i & 1;
        int fr, start, npts;
        tile_info(tile_of(i), fr, start, npts);
        tr.mark(0x10);
#pragma unroll
        for (int l = 0; l < S::NH; ++l) {
          const uint32_t a_buf = sbase + L::buf_off(S::ACT_BUF(l));
          const int nbn = (S::HN(l) + 127) / 128, kbn = S::HK(l) / 64;
          for (int sub = 0; sub < NSUB; ++sub) {
            const int r = hidden_region(sub);
            mbar_wait_w(act_ready(l, sub), tpar);
            tr.mark(0x20 + l * 2 + sub);
            if (S::EMIT_LAYER >= 0 && l == S::EMIT_LAYER + 1 && args.emit != nullptr) {
              // point_feat of this sub-tile: the swizzled [128 x 64] bf16 operand image goes to HBM as is
              const size_t trow = ((size_t)fr * tiles_per_frustum + start / TILE) * TILE + sub * 128;
              if (lane == 0) bulk_s2g(reinterpret_cast<uint8_t*>(args.emit) + trow * 128, a_buf + sub * (128 * 128), 128 * 128);
              __syncwarp();
            }
            mbar_wait_w(acc_empty(r), (acc_cnt[r] & 1) ^ 1);
            tc_fence_after();
            for (int nb = 0; nb < nbn; ++nb) {
              const int ncols = min(128, S::HN(l) - nb * 128);
              const uint32_t idesc = make_idesc_bf16(128, ncols);
              for (int kb = 0; kb < kbn; ++kb) {
                const uint32_t itc = it + nb * kbn + kb;
                const int s = itc % kRingStages;
                if (sub == 0) { mbar_wait_w(ring_full(s), (itc / kRingStages) & 1); tc_fence_after(); }
                const uint32_t b_addr = sbase + L::RING + s * kChunkBytes;
                const uint32_t a_addr = a_buf + kb * (TILE * 128) + sub * (128 * 128);
                const uint32_t d = tmem_base + region_col(r) + nb * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16_w(d, make_sdesc_k128(a_addr + k * 32), make_sdesc_k128(b_addr + k * 32), idesc, (kb | k) != 0);
                if (sub == NSUB - 1) umma_commit_mc_w(ring_empty(s), kAllCtas);
              }
            }
            umma_commit_w(acc_full(r));
            acc_cnt[r]++;
            tr.mark(0x30 + l * 2 + sub);
          }
          it += nbn * kbn;
          if (l == S::FRONT_FREE_LAYER) {
            if (S::EMIT_LAYER >= 0 && args.emit != nullptr) { if (lane == 0) bulk_wait_read_all(); __syncwarp(); }   // emit has left the smem buffer
            umma_commit_w(front_free);
          }
        }
        // final layer: channels on M, points on N
        for (int sub = 0; sub < NSUB; ++sub) mbar_wait_w(act_ready(S::NH, sub), tpar);
        tr.mark(0x40);
        tc_fence_after();
        const uint32_t b_buf = sbase + L::buf_off(S::ACT_BUF(S::NH));
        const uint32_t idesc_f = make_idesc_bf16(128, TILE);
        for (int mt = 0; mt < NMT; ++mt) {
          const int r = mt & 1;
          mbar_wait_w(acc_empty(r), (acc_cnt[r] & 1) ^ 1);
          tr.mark(0x50 + mt);
          tc_fence_after();
          const uint32_t d = tmem_base + region_col(r);
          for (int kb = 0; kb < S::FK / 64; ++kb, ++it) {
            const int s = it % kRingStages;
            mbar_wait_w(ring_full(s), (it / kRingStages) & 1);
            tc_fence_after();
            const uint32_t a_addr = sbase + L::RING + s * kChunkBytes;
            const uint32_t b_addr = b_buf + kb * (TILE * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_w(d, make_sdesc_k128(a_addr + k * 32), make_sdesc_k128(b_addr + k * 32), idesc_f, (kb | k) != 0);
            umma_commit_mc_w(ring_empty(s), kAllCtas);
          }
          umma_commit_w(acc_full(r));
          acc_cnt[r]++;
          tr.mark(0x60 + mt);
        }
        if (S::NH == 0) umma_commit_w(front_free);
      }
      if (S::EMIT_LAYER >= 0 && args.emit != nullptr && lane == 0) bulk_wait_all();
    }
  } else if (warp >= 4 && warp < 12) {
    // ================================================================ epilogue warps: TMEM lane quarter = warp&3, column half = (warp-4)>>2
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q * 32 + lane;                 // TMEM lane owned by this thread
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    uint32_t acc_cnt[3] = {0, 0, 0};
    float run_max[NMT];
#pragma unroll
    for (int i = 0; i < NMT; ++i) run_max[i] = -3.0e38f;
    const uint32_t hbias = sbase + L::HB;            // fp32 hidden biases in smem (byte address)
    const float* fbias = reinterpret_cast<const float*>(args.arena + (size_t)chain_num_chunks<S>() * kChunkBytes) +
                         S::CIN * S::C1 + S::C1 + chain_hidden_bias_count<S>();
    Tracer tr; tr.init((warp == 4 && lane == 0) ? args.trace : nullptr, 2);

    // 32 accumulator columns [c0, c0+32) of this thread's row: +bias, ReLU, bf16, into the next operand
    auto store_group = [&](const uint32_t (&v)[32], uint32_t bias, int c0, uint32_t o_buf, uint32_t grow) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = ld_shared_f4(bias + 4u * (c0 + 4 * j));
        pk[2 * j] = pack_bf16_relu(__uint_as_float(v[4 * j]) + b4.x, __uint_as_float(v[4 * j + 1]) + b4.y);
        pk[2 * j + 1] = pack_bf16_relu(__uint_as_float(v[4 * j + 2]) + b4.z, __uint_as_float(v[4 * j + 3]) + b4.w);
      }
      const int kb = c0 >> 6, j0 = (c0 & 63) >> 3;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        st_shared_v4(o_buf + kb * (TILE * 128) + sw128_offset(grow, j0 + jj), pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
    };

    for (int i = 0; i < iters; ++i) {
      int fr, start, npts;
      tile_info(tile_of(i), fr, start, npts);
      int hb_off = 0;
      tr.mark(0x10);
#pragma unroll
      for (int l = 0; l < S::NH; ++l) {
        const uint32_t o_buf = sbase + L::buf_off(S::ACT_BUF(l + 1));
        const int span = S::HN(l) / 2, cbeg = half * span;
        for (int sub = 0; sub < NSUB; ++sub) {
          const int r = hidden_region(sub);
          mbar_wait(acc_full(r), acc_cnt[r] & 1);
          acc_cnt[r]++;
          tr.mark(0x20 + l * 2 + sub);
          tc_fence_after();
          const uint32_t grow = sub * 128 + row;         // row inside the tile
          const uint32_t taddr = tmem_base + lane_sel + region_col(r);
          if (span == 32) {
            uint32_t v[32];
            tmem_ld32(taddr + cbeg, v);
            tmem_ld_wait();
            store_group(v, hbias + 4u * hb_off, cbeg, o_buf, grow);
          } else {
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + span; c0 += 64) {
              uint32_t va[32], vb[32];
              tmem_ld32(taddr + c0, va);
              tmem_ld32(taddr + c0 + 32, vb);
              tmem_ld_wait();
              store_group(va, hbias + 4u * hb_off, c0, o_buf, grow);
              store_group(vb, hbias + 4u * hb_off, c0 + 32, o_buf, grow);
            }
          }
          tc_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { mbar_arrive(act_ready(l + 1, sub)); mbar_arrive(acc_empty(r)); }
          tr.mark(0x30 + l * 2 + sub);
        }
        hb_off += S::HN(l);
      }
      // final layer: this thread owns channel mt*128+row, columns [half*TILE/2, +TILE/2) are points
#pragma unroll
      for (int mt = 0; mt < NMT; ++mt) {
        const int r = mt & 1;
        mbar_wait(acc_full(r), acc_cnt[r] & 1);
        acc_cnt[r]++;
        tr.mark(0x50 + mt);
        tc_fence_after();
        float m = run_max[mt];
        const uint32_t taddr = tmem_base + lane_sel + region_col(r) + half * (TILE / 2);
#pragma unroll 1
        for (int c0 = 0; c0 < TILE / 2; c0 += 64) {
          uint32_t va[32], vb[32];
          tmem_ld32(taddr + c0, va);
          tmem_ld32(taddr + c0 + 32, vb);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) m = fmax3(m, __uint_as_float(va[2 * j]), __uint_as_float(va[2 * j + 1]));
#pragma unroll
          for (int j = 0; j < 16; ++j) m = fmax3(m, __uint_as_float(vb[2 * j]), __uint_as_float(vb[2 * j + 1]));
        }
        run_max[mt] = m;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(r));
        tr.mark(0x60 + mt);
      }
      // flush when the frustum changes (or at the end of this CTA's range)
      bool flush = (i + 1 == iters);
      if (!flush) { int f2, s2, n2; tile_info(tile_of(i + 1), f2, s2, n2); flush = (f2 != fr); }
      if (flush) {
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt) {
          const int ch = mt * 128 + row;
          const float v = fmaxf(run_max[mt] + fbias[ch], 0.0f);
          atomicMax(reinterpret_cast<int*>(args.out + (size_t)fr * S::FC + ch), __float_as_int(v));
          run_max[mt] = -3.0e38f;
        }
      }
    }
  } else if (warp >= 12) {
    // ================================================================ front warps: load points, layer 1 on CUDA cores
    const int p = threadIdx.x - 384;               // 0..127
    const uint32_t W1 = sbase + L::W1, B1 = sbase + L::B1;      // fp32 layer-1 weights / bias in smem (byte addresses)
    const uint32_t o_buf = sbase + L::buf_off(S::ACT_BUF(0));
    Tracer tr; tr.init((warp == 12 && lane == 0) ? args.trace : nullptr, 3);
    for (int i = 0; i < iters; ++i) {
      int fr, start, npts;
      tile_info(tile_of(i), fr, start, npts);
      if (i > 0) mbar_wait(front_free, (i - 1) & 1);
      tr.mark(0x10);
      float cx = 0.f, cy = 0.f, cz = 0.f;
      if (args.center) { cx = args.center[fr * 3 + 0]; cy = args.center[fr * 3 + 1]; cz = args.center[fr * 3 + 2]; }
      float bc[3] = {0, 0, 0}, hl = 0, hw = 0, hh = 0, ct = 1, st = 0;
      if (S::BOXPC) {
        bc[0] = args.box_center[fr * 3 + 0]; bc[1] = args.box_center[fr * 3 + 1]; bc[2] = args.box_center[fr * 3 + 2];
        hl = 0.5f * args.box_dims[fr * 3 + 0]; hw = 0.5f * args.box_dims[fr * 3 + 1]; hh = 0.5f * args.box_dims[fr * 3 + 2];
        sincosf(args.box_orient[fr], &st, &ct);
      }
      for (int sub = 0; sub < NSUB; ++sub) {
        const int grow = sub * 128 + p;
        int j = min(grow, npts - 1);               // padded rows duplicate the last valid point
        j += start;
        const int src = args.idx ? args.idx[(size_t)fr * args.idx_stride + j] : j;
        float x[S::CIN];
        if (S::CRAW == 6 && args.rgb != nullptr) {
          const size_t pi = ((size_t)fr * args.N + src) * 3;
#pragma unroll
          for (int k = 0; k < 3; ++k) { x[k] = args.pc[pi + k]; x[3 + k] = __fdiv_rn((float)args.rgb[pi + k], 255.0f); }
        } else {
          const float* pp = args.pc + ((size_t)fr * args.N + src) * args.C;
#pragma unroll
          for (int k = 0; k < S::CRAW; ++k) x[k] = pp[k];
        }
        x[0] -= cx; x[1] -= cy; x[2] -= cz;
        if (S::BOXPC) {
          // 6 signed plane distances (models/tf_util.py:764-795, closed form in SURVEY a14)
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
          st_shared_v4(o_buf + (c0 >> 6) * (TILE * 128) + sw128_offset(grow, (c0 & 63) >> 3),
                       pack_bf16_relu(a[0], a[1]), pack_bf16_relu(a[2], a[3]), pack_bf16_relu(a[4], a[5]), pack_bf16_relu(a[6], a[7]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(act_ready(0, sub));
      }
      tr.mark(0x11);
    }
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();              // no CTA exits while its peer may still multicast into it
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

}  // namespace t3d
