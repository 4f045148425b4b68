// CUDA-core kernels of the hot path: the fp32 layer GEMM with fused bias / per-frustum bias /
// activation / row-mask / max-over-points epilogue (fp32 mode of every conv2d / fully_connected,
// models/tf_util.py:1258-1323,1463-1499, and the FC heads in both modes), the mask / centroid /
// compaction kernel (semisup_models.py:145-162, model_util.py:241-272), Philox resampling + gather
// (model_util.py:61-91), BoxPC features (tf_util.py:764-795), output parsing + anchor->reg
// (semisup_models.py:265-290, tf_util.py:1001-1041), BoxPC refine (test_semisup.py:116-142),
// box corners (model_util.py:94-167) and weight packing for the tcgen05 kernels.
#pragma once
#include "common.cuh"

namespace t3d {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_TANH = 3 };

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.0f);
    case ACT_LEAKY: return v > 0.0f ? v : 0.2f * v;         // tf.nn.leaky_relu default alpha
    case ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// atomic max on floats; requires the target to be initialised to 0 and is used on values that are
// >= 0 after ReLU / mask (negative values fall back to the ordered-uint trick for completeness).
__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

struct LinearArgs {
  const float* X; int ldx;
  const float* W; int ldw;            // [K, N] row-major (TF layout)
  const float* bias;                  // [N] or null
  const float* gbias;                 // [groups, N] or null, group = row / rows_per_group
  int rows_per_group;
  float* Y; int ldy;                  // null when only the group max is wanted
  int M, K, N, act;
  const float* rowmask;               // [M] multiplies the activated row (net*mask) or null
  float* gmax;                        // [groups, N] zero-initialised: max over the rows of each group
};

__global__ void __launch_bounds__(256) linear_f32_kernel(const LinearArgs a) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN];
  __shared__ float red[16][BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < a.K; k0 += BK) {
    {
      const int r = tid >> 2, kk = (tid & 3) * 4;
      const int gm = m0 + r;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int gk = k0 + kk + e;
        As[kk + e][r] = (gm < a.M && gk < a.K) ? a.X[(size_t)gm * a.ldx + gk] : 0.0f;
      }
      const int k = tid >> 4, nn = (tid & 15) * 4;
      const int gk = k0 + k;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int gn = n0 + nn + e;
        Bs[k][nn + e] = (gk < a.K && gn < a.N) ? a.W[(size_t)gk * a.ldw + gn] : 0.0f;
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
  // epilogue
  const bool one_group = a.gmax && (a.rows_per_group % BM == 0) && (m0 + BM <= a.M);
  float cmax[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= a.M) continue;
    const int g = a.rows_per_group > 0 ? gm / a.rows_per_group : 0;
    const float rm = a.rowmask ? a.rowmask[gm] : 1.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= a.N) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[gn];
      if (a.gbias) v += a.gbias[(size_t)g * a.N + gn];
      v = apply_act(v, a.act) * rm;
      if (a.Y) a.Y[(size_t)gm * a.ldy + gn] = v;
      if (a.gmax) {
        if (one_group) cmax[j] = fmaxf(cmax[j], v);
        else atomic_max_f32(a.gmax + (size_t)g * a.N + gn, v);
      }
    }
  }
  if (one_group) {   // values are >= 0 here (ReLU / mask), reduce the tile's 64 rows before the atomics
#pragma unroll
    for (int j = 0; j < 4; ++j) red[ty][tx * 4 + j] = cmax[j];
    __syncthreads();
    if (tid < BN) {
      float m = red[0][tid];
#pragma unroll
      for (int r = 1; r < 16; ++r) m = fmaxf(m, red[r][tid]);
      const int gn = n0 + tid;
      if (gn < a.N) atomic_max_f32(a.gmax + (size_t)(m0 / a.rows_per_group) * a.N + gn, m);
    }
  }
}

// ----------------------------------------------------------------------------- mask + centroid + compaction
// One CTA per frustum. mask = float(l0 < l1) (strict, ties -> 0); mean = sum(mask*xyz)/max(count,1);
// idx = ascending list of masked-in point indices (ballot + prefix sum), count = its length.
__global__ void __launch_bounds__(256) mask_centroid_kernel(const float* __restrict__ logits, const float* __restrict__ pc,
                                                            int N, int C, float* __restrict__ mask, int* __restrict__ count,
                                                            float* __restrict__ mean, float* __restrict__ xyz_stage1,
                                                            int* __restrict__ idx) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int wcnt[8];
  __shared__ float wsum[8][3];
  __shared__ int base_s;
  __shared__ float mean_s[3];
  if (tid == 0) base_s = 0;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  __syncthreads();
  for (int p0 = 0; p0 < N; p0 += 256) {
    const int p = p0 + tid;
    bool m = false;
    if (p < N) {
      const float2 l = *reinterpret_cast<const float2*>(logits + ((size_t)b * N + p) * 2);
      m = l.x < l.y;
      if (mask) mask[(size_t)b * N + p] = m ? 1.0f : 0.0f;
      if (m) {
        const float* q = pc + ((size_t)b * N + p) * C;
        sx += q[0]; sy += q[1]; sz += q[2];
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    if (lane == 0) wcnt[warp] = __popc(bal);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < warp; ++w) off += wcnt[w];
    if (m && idx) idx[(size_t)b * N + off + __popc(bal & ((1u << lane) - 1))] = p;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += wcnt[w]; base_s += t; }
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sz += __shfl_xor_sync(0xffffffffu, sz, o);
  }
  if (lane == 0) { wsum[warp][0] = sx; wsum[warp][1] = sy; wsum[warp][2] = sz; }
  __syncthreads();
  if (tid == 0) {
    float t[3] = {0, 0, 0};
    for (int w = 0; w < 8; ++w) { t[0] += wsum[w][0]; t[1] += wsum[w][1]; t[2] += wsum[w][2]; }
    const int n = base_s;
    const float d = fmaxf((float)n, 1.0f);
    for (int k = 0; k < 3; ++k) { mean_s[k] = t[k] / d; if (mean) mean[b * 3 + k] = mean_s[k]; }
    if (count) count[b] = n;
  }
  __syncthreads();
  if (xyz_stage1) {
    for (int p = tid; p < N; p += 256) {
      const float* q = pc + ((size_t)b * N + p) * C;
      float* o = xyz_stage1 + ((size_t)b * N + p) * 3;
      o[0] = q[0] - mean_s[0]; o[1] = q[1] - mean_s[1]; o[2] = q[2] - mean_s[2];
    }
  }
}

// tile table for the ragged (compacted) tcgen05 chains: tiles[t] = {frustum, start, npts, 0}
__global__ void build_tiles_kernel(const int* __restrict__ count, int B, int tile_pts, int4* __restrict__ tiles,
                                   int* __restrict__ num_tiles) {
  __shared__ int carry;
  __shared__ int scan[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    const int n = b < B ? count[b] : 0;
    const int nt = (n + tile_pts - 1) / tile_pts;
    scan[threadIdx.x] = nt;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int v = threadIdx.x >= o ? scan[threadIdx.x - o] : 0;
      __syncthreads();
      scan[threadIdx.x] += v;
      __syncthreads();
    }
    const int first = carry + scan[threadIdx.x] - nt;
    for (int i = 0; i < nt; ++i) tiles[first + i] = make_int4(b, i * tile_pts, min(tile_pts, n - i * tile_pts), 0);
    __syncthreads();
    if (threadIdx.x == 1023) carry += scan[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_tiles = carry;
}

// ----------------------------------------------------------------------------- Philox4x32-10 resampling
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&o)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// One CTA per frustum; supports count <= 2048 and npoints <= 2048.
// mode 0 ('philox'): see oracle/model_util.py philox_choice -- random ordered subset when count > npoints,
// identity ++ uniform refill then a random shuffle otherwise; selection = stable sort by 64-bit key.
// mode 1 ('choice'): rank-space choice array supplied by the host (numpy_legacy stream).
// indices[b,t] = {b, point}; object_pc[b,t,:] = (xyz - mean, features) of that point.
//
// The sort is a one-pass bucket sort: the keys are uniform 64-bit randoms, so their top 9 bits spread the <= 2048
// elements over 512 buckets of ~4; an element's final rank is its bucket's start (histogram + scan) plus the number of
// bucket mates that order before it under the (key, element id) comparison, i.e. exactly the stable-argsort position.
// Only ranks < npoints are materialised.  ~12 shared-memory accesses per element and 5 block barriers instead of the
// 66 compare-exchange stages of a 2048-wide bitonic network.
constexpr int kResampleThreads = 512;
constexpr int kResampleBuckets = 512;
__global__ void __launch_bounds__(kResampleThreads) resample_kernel(const int* __restrict__ idx, const int* __restrict__ count, int N, int npoints,
                                                        int mode, unsigned long long seed, const int* __restrict__ choice,
                                                        int* __restrict__ indices, const float* __restrict__ pc, int C,
                                                        const float* __restrict__ mean, int c_out, float* __restrict__ object_pc) {
  __shared__ unsigned long long keys[2048];
  __shared__ unsigned short lst[2048];
  __shared__ unsigned short order[2048];     // element ids grouped by bucket
  __shared__ unsigned short sel[2048];       // sel[rank] = element id, rank < npoints
  __shared__ int start[kResampleBuckets + 1];
  __shared__ int cursor[kResampleBuckets];
  __shared__ int warp_tot[kResampleThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = count[b];
  const uint32_t k0 = (uint32_t)(seed & 0xFFFFFFFFull), k1 = (uint32_t)(seed >> 32);
  if (n > 0 && mode == 0) {
    const bool sub = n > npoints;
    const int len = sub ? n : npoints;
    for (int t = tid; t < kResampleBuckets; t += kResampleThreads) cursor[t] = 0;
    __syncthreads();
    for (int t = tid; t < len; t += kResampleThreads) {
      uint32_t o[4];
      philox4x32_10((uint32_t)t, sub ? 0u : 2u, (uint32_t)b, 0u, k0, k1, o);
      const unsigned long long key = ((unsigned long long)o[0] << 32) | o[1];
      if (!sub) {
        uint32_t r[4];
        philox4x32_10((uint32_t)t, 1u, (uint32_t)b, 0u, k0, k1, r);
        lst[t] = (unsigned short)(t < n ? t : (r[2] % (uint32_t)n));
      }
      keys[t] = key;
      atomicAdd(&cursor[(int)(key >> 55)], 1);
    }
    __syncthreads();
    {   // exclusive scan of the bucket counts (one bucket per thread)
      const int c = cursor[tid];
      int v = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, d); if ((tid & 31) >= d) v += u; }
      if ((tid & 31) == 31) warp_tot[tid >> 5] = v;
      __syncthreads();
      int base = 0;
      for (int w = 0; w < (tid >> 5); ++w) base += warp_tot[w];
      start[tid] = base + v - c;
      if (tid == kResampleThreads - 1) start[kResampleBuckets] = base + v;
      cursor[tid] = base + v - c;
    }
    __syncthreads();
    for (int t = tid; t < len; t += kResampleThreads) {
      const int pos = atomicAdd(&cursor[(int)(keys[t] >> 55)], 1);
      order[pos] = (unsigned short)t;
    }
    __syncthreads();
    for (int t = tid; t < len; t += kResampleThreads) {
      const unsigned long long key = keys[t];
      const int bk = (int)(key >> 55);
      const int s = start[bk], e = start[bk + 1];
      if (s >= npoints) continue;                    // the whole bucket ranks beyond what is kept
      int r = s;
      for (int u = s; u < e; ++u) {
        const int v = order[u];
        const unsigned long long kv = keys[v];
        r += (kv < key) || (kv == key && v < t);     // ties by element id (== numpy stable argsort)
      }
      if (r < npoints) sel[r] = (unsigned short)t;
    }
    __syncthreads();
  }
  for (int t = tid; t < npoints; t += kResampleThreads) {
    int point = 0;
    if (n > 0) {
      int rank;
      if (mode == 0) rank = (n > npoints) ? sel[t] : lst[sel[t]];
      else rank = choice[(size_t)b * npoints + t];
      point = idx[(size_t)b * N + rank];
    }
    indices[((size_t)b * npoints + t) * 2 + 0] = b;
    indices[((size_t)b * npoints + t) * 2 + 1] = point;
    if (object_pc) {
      const float* q = pc + ((size_t)b * N + point) * C;
      float* o = object_pc + ((size_t)b * npoints + t) * c_out;
      for (int k = 0; k < c_out; ++k) o[k] = k < 3 ? q[k] - mean[b * 3 + k] : q[k];
    }
  }
}

// ----------------------------------------------------------------------------- fp32-mode input builders
// out[b,p,0:3] = xyz - center[b] (center may be null); optional gather through idx.
__global__ void prepare_xyz_kernel(const float* __restrict__ pc, int B, int N, int C, const float* __restrict__ center,
                                   float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * N) return;
  const int b = (int)(i / N);
  const float* q = pc + i * C;
  float cx = 0, cy = 0, cz = 0;
  if (center) { cx = center[b * 3]; cy = center[b * 3 + 1]; cz = center[b * 3 + 2]; }
  out[i * 3 + 0] = q[0] - cx; out[i * 3 + 1] = q[1] - cy; out[i * 3 + 2] = q[2] - cz;
}

// BoxPC representation (tf_util.py:764-795): out[b,p,:] = [pc (C channels, untranslated), 6 plane distances]
__global__ void boxpc_features_kernel(const float* __restrict__ pc, int B, int N, int C, const float* __restrict__ center,
                                      const float* __restrict__ dims, const float* __restrict__ orient, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * N) return;
  const int b = (int)(i / N);
  const float* q = pc + i * C;
  float* o = out + i * (C + 6);
  for (int k = 0; k < C; ++k) o[k] = q[k];
  float st, ct;
  sincosf(orient[b], &st, &ct);
  const float dx = q[0] - center[b * 3], dy = q[1] - center[b * 3 + 1], dz = q[2] - center[b * 3 + 2];
  const float hl = 0.5f * dims[b * 3], hw = 0.5f * dims[b * 3 + 1], hh = 0.5f * dims[b * 3 + 2];
  const float xr = ct * dx - st * dz, zr = st * dx + ct * dz;
  o[C + 0] = hl - xr; o[C + 1] = hl + xr; o[C + 2] = hh - dy; o[C + 3] = hh + dy; o[C + 4] = hw - zr; o[C + 5] = hw + zr;
}

// ----------------------------------------------------------------------------- heads
// parse (semisup_models.py:265-290 / semisup_v1_sunrgbd.py:203-222 / model_util.py:178-210) fused with
// anchor->reg (tf_util.py:1001-1041). One thread per frustum. Any output pointer may be null.
struct ParseArgs {
  const float* output;         // [B, 3+2NH+4NS]
  const float* stage1_center;  // [B,3] added to the centre (null: no add, = parse_output_to_tensors)
  const float* mean_size;      // [NS,3] dims anchors
  const float* orient_anchors; // [NH] heading bin centres
  int B, NH, NS;
  float *center, *heading_scores, *heading_res_norm, *heading_res, *size_scores, *size_res_norm, *size_res;
  float *reg_center, *reg_dims, *reg_orient;   // box in (center, dims, orient) regression format
};
__global__ void parse_box_kernel(const ParseArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const int NH = a.NH, NS = a.NS, W = 3 + 2 * NH + 4 * NS;
  const float* o = a.output + (size_t)b * W;
  float c[3];
  for (int k = 0; k < 3; ++k) { c[k] = o[k] + (a.stage1_center ? a.stage1_center[b * 3 + k] : 0.0f); if (a.center) a.center[b * 3 + k] = c[k]; }
  const float hscale = 3.14159265358979323846f / (float)NH;
  int hbest = 0; float hmax = o[3];
  for (int j = 0; j < NH; ++j) {
    const float s = o[3 + j], rn = o[3 + NH + j];
    if (s > hmax) { hmax = s; hbest = j; }                // first max wins (tf.argmax)
    if (a.heading_scores) a.heading_scores[b * NH + j] = s;
    if (a.heading_res_norm) a.heading_res_norm[b * NH + j] = rn;
    if (a.heading_res) a.heading_res[b * NH + j] = rn * hscale;
  }
  int sbest = 0; float smax = o[3 + 2 * NH];
  for (int j = 0; j < NS; ++j) {
    const float s = o[3 + 2 * NH + j];
    if (s > smax) { smax = s; sbest = j; }
    if (a.size_scores) a.size_scores[b * NS + j] = s;
    for (int k = 0; k < 3; ++k) {
      const float rn = o[3 + 2 * NH + NS + j * 3 + k];
      if (a.size_res_norm) a.size_res_norm[(b * NS + j) * 3 + k] = rn;
      if (a.size_res) a.size_res[(b * NS + j) * 3 + k] = rn * a.mean_size[j * 3 + k];
    }
  }
  if (a.reg_center) {
    for (int k = 0; k < 3; ++k) {
      a.reg_center[b * 3 + k] = c[k];
      const float d = a.mean_size[sbest * 3 + k] + o[3 + 2 * NH + NS + sbest * 3 + k] * a.mean_size[sbest * 3 + k];
      a.reg_dims[b * 3 + k] = fmaxf(d, 1e-5f);
    }
    a.reg_orient[b] = a.orient_anchors[hbest] + o[3 + NH + hbest] * hscale;
  }
}

// generic anchor->reg on already-parsed tensors (tf_util.py:1001-1041; used for label boxes too)
__global__ void anchor_to_reg_kernel(const float* center, const float* dims_cls, const float* dims_reg, const float* orient_cls,
                                     const float* orient_reg, const float* dims_anchors, const float* orient_anchors,
                                     int B, int NS, int NH, float* out_center, float* out_dims, float* out_orient) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int i = 0; float m = dims_cls[b * NS];
  for (int j = 1; j < NS; ++j) if (dims_cls[b * NS + j] > m) { m = dims_cls[b * NS + j]; i = j; }
  int h = 0; m = orient_cls[b * NH];
  for (int j = 1; j < NH; ++j) if (orient_cls[b * NH + j] > m) { m = orient_cls[b * NH + j]; h = j; }
  for (int k = 0; k < 3; ++k) {
    out_center[b * 3 + k] = center[b * 3 + k];
    out_dims[b * 3 + k] = fmaxf(dims_anchors[i * 3 + k] + dims_reg[(b * NS + i) * 3 + k], 1e-5f);
  }
  out_orient[b] = orient_anchors[h] + orient_reg[b * NH + h];
}

// BoxPC output parsing + one refine step (boxpc_sunrgbd.py:70-95, test_semisup.py:116-134).
struct RefineArgs {
  const float* out9;     // [B,9]: dc(3), ds(3), da(1), fit logits(2)
  int B;
  int weigh_pred_by_conf;    // c.BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF
  int weigh_during_test;     // FLAGS.SEMI_WEIGH_BOXPC_DELTA_DURING_TEST
  float *fit_logits, *fit_prob; int* pred_fit;
  float *delta_center, *delta_size, *delta_angle;     // model deltas (end_points boxpc_delta_*)
  float *box_center, *box_dims, *box_orient;          // curr_box, updated in place (may be null)
  float *tot_center, *tot_size, *tot_angle;           // accumulated deltas, updated in place (may be null)
};
__global__ void boxpc_refine_kernel(const RefineArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const float* o = a.out9 + (size_t)b * 9;
  const float l0 = o[7], l1 = o[8];
  const float mx = fmaxf(l0, l1);
  const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
  const float p1 = e1 / (e0 + e1);
  if (a.fit_logits) { a.fit_logits[b * 2] = l0; a.fit_logits[b * 2 + 1] = l1; }
  if (a.fit_prob) a.fit_prob[b] = p1;
  if (a.pred_fit) a.pred_fit[b] = p1 > 0.5f ? 1 : 0;
  const float wp = a.weigh_pred_by_conf ? (1.0f - p1) : 1.0f;
  float dc[3], ds[3], da = o[6] * wp;
  for (int k = 0; k < 3; ++k) { dc[k] = o[k] * wp; ds[k] = o[3 + k] * wp; }
  if (a.delta_center) for (int k = 0; k < 3; ++k) { a.delta_center[b * 3 + k] = dc[k]; a.delta_size[b * 3 + k] = ds[k]; }
  if (a.delta_angle) a.delta_angle[b] = da;
  const float wt = a.weigh_during_test ? (1.0f - p1) : 1.0f;
  if (a.box_center) {
    for (int k = 0; k < 3; ++k) { a.box_center[b * 3 + k] -= dc[k] * wt; a.box_dims[b * 3 + k] -= ds[k] * wt; }
    a.box_orient[b] -= da * wt;
  }
  if (a.tot_center) {
    for (int k = 0; k < 3; ++k) { a.tot_center[b * 3 + k] += dc[k] * wt; a.tot_size[b * 3 + k] += ds[k] * wt; }
    a.tot_angle[b] += da * wt;
  }
}

// F2_* = F_* - total deltas (test_semisup.py:136-142)
__global__ void f2_kernel(const float* f_center, const float* f_hres, const float* f_sres, const float* tot_center,
                          const float* tot_angle, const float* tot_size, int B, int NH, int NS,
                          float* f2_center, float* f2_hres, float* f2_sres) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int k = 0; k < 3; ++k) f2_center[b * 3 + k] = f_center[b * 3 + k] - tot_center[b * 3 + k];
  for (int j = 0; j < NH; ++j) f2_hres[b * NH + j] = f_hres[b * NH + j] - tot_angle[b];
  for (int j = 0; j < NS; ++j)
    for (int k = 0; k < 3; ++k) f2_sres[(b * NS + j) * 3 + k] = f_sres[(b * NS + j) * 3 + k] - tot_size[b * 3 + k];
}

// get_box3d_corners_helper (model_util.py:94-119): n boxes -> (n,8,3)
__device__ __forceinline__ void box_corners(const float c[3], float heading, const float s[3], float* out /*8x3*/) {
  const float l = s[0], w = s[1], h = s[2];
  const float xs[8] = {1, 1, -1, -1, 1, 1, -1, -1}, ys[8] = {1, 1, 1, 1, -1, -1, -1, -1}, zs[8] = {1, -1, -1, 1, 1, -1, -1, 1};
  float sn, cs;
  sincosf(heading, &sn, &cs);
  for (int i = 0; i < 8; ++i) {
    const float x = xs[i] * l * 0.5f, y = ys[i] * h * 0.5f, z = zs[i] * w * 0.5f;
    out[i * 3 + 0] = cs * x + sn * z + c[0];
    out[i * 3 + 1] = y + c[1];
    out[i * 3 + 2] = -sn * x + cs * z + c[2];
  }
}
__global__ void box3d_corners_helper_kernel(const float* centers, const float* headings, const float* sizes, int n, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  box_corners(centers + i * 3, headings[i], sizes + i * 3, out + (size_t)i * 24);
}
// get_box3d_corners(_sunrgbd) (model_util.py:121-167): all NH x NS candidate boxes; sizes = mean + 2*residual
__global__ void box3d_corners_all_kernel(const float* center, const float* heading_res, const float* size_res,
                                         const float* mean_size, const float* orient_anchors, int B, int NH, int NS, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * NH * NS) return;
  const int s = i % NS, h = (i / NS) % NH, b = i / (NS * NH);
  const float heading = heading_res[b * NH + h] + orient_anchors[h];
  float sz[3];
  for (int k = 0; k < 3; ++k) { const float r = size_res[(b * NS + s) * 3 + k]; sz[k] = (mean_size[s * 3 + k] + r) + r; }
  box_corners(center + b * 3, heading, sz, out + (size_t)i * 24);
}

// ----------------------------------------------------------------------------- point-cloud normalisation (NORMALIZE_PC options)
// tf_normalize_point_clouds_to_mean_zero_and_unit_var (mode 0, models/tf_util.py:157-173): (x - mean) / (sqrt(var) + 1e-5) per
// cloud and channel, biased variance; tf_normalize_point_clouds_to_01 (mode 1, :134-155): (x - mean) / (largest xyz extent
// + 1e-5).  Only the first 3 channels are normalised, channels >= 3 are copied (the concat of both reference functions).
// One CTA per cloud; mean first, then the centred second pass (what tf.nn.moments computes), then the write.
__device__ __forceinline__ float block_reduce(float v, int op, float* sh) {      // op 0 sum, 1 max, 2 min; 256 threads
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = op == 0 ? v + w : (op == 1 ? fmaxf(v, w) : fminf(v, w));
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int i = 1; i < 8; ++i) r = op == 0 ? r + sh[i] : (op == 1 ? fmaxf(r, sh[i]) : fminf(r, sh[i]));
  return r;
}
__global__ void __launch_bounds__(256) normalize_pc_kernel(const float* __restrict__ pc, int N, int C, int mode, float* __restrict__ out) {
  __shared__ float sh[8];
  const float* p = pc + (size_t)blockIdx.x * N * C;
  float* q = out + (size_t)blockIdx.x * N * C;
  float mean[3], div[3];
  for (int k = 0; k < 3; ++k) {
    float s = 0.f;
    for (int n = threadIdx.x; n < N; n += 256) s += p[(size_t)n * C + k];
    mean[k] = block_reduce(s, 0, sh) / (float)N;
  }
  if (mode == 0) {
    for (int k = 0; k < 3; ++k) {
      float s = 0.f;
      for (int n = threadIdx.x; n < N; n += 256) { const float d = p[(size_t)n * C + k] - mean[k]; s = fmaf(d, d, s); }
      div[k] = sqrtf(block_reduce(s, 0, sh) / (float)N) + 1e-5f;
    }
  } else {
    float ext = -INFINITY;
    for (int k = 0; k < 3; ++k) {
      float mx = -INFINITY, mn = INFINITY;
      for (int n = threadIdx.x; n < N; n += 256) { const float d = p[(size_t)n * C + k] - mean[k]; mx = fmaxf(mx, d); mn = fminf(mn, d); }
      ext = fmaxf(ext, block_reduce(mx, 1, sh) - block_reduce(mn, 2, sh));
    }
    div[0] = div[1] = div[2] = ext + 1e-5f;
  }
  for (int n = threadIdx.x; n < N; n += 256) {
    for (int k = 0; k < 3; ++k) q[(size_t)n * C + k] = (p[(size_t)n * C + k] - mean[k]) / div[k];
    for (int k = 3; k < C; ++k) q[(size_t)n * C + k] = p[(size_t)n * C + k];
  }
}

// ----------------------------------------------------------------------------- weight packing for tcgen05
// chunk image: [128 rows x 64 K] bf16, K-major SWIZZLE_128B; row r = output channel row0+r, K = k0..k0+63 of
// W[K_total, Nout] (row-major, TF layout).  Rows >= nrows and k >= K_total are zero.
// one 16 KB weight chunk image: rows [row0, row0+nrows) x K [k0, k0+64) of W[k_total, ldw]; part 0 = bf16,
// 1 / 2 = fp16 hi / lo image of W * scale (f16x2 kernels)
struct PackDesc { const float* W; int ldw; int k_total; int k0; int row0; int nrows; int part; float scale; };

}  // namespace t3d
