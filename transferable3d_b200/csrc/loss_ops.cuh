// Loss kernels of the semi-supervised training step (train_semisup_adv.py:267-425, BASELINE cfg5):
//   * strong loss      semisup_v1_sunrgbd.get_strong_loss          (semisup_v1_sunrgbd.py:423-553)
//   * reprojection     weak_losses.get_reprojection_loss           (weak_losses.py:69-238)
//   * intra-class var  weak_losses.get_intraclass_variance_loss_v1 (weak_losses.py:267-291)
//   * BoxPC fit loss   -log(0.01 + p_fit)                          (semisup_v1_sunrgbd.py:399-411)
//   * combination      semisup_v1_sunrgbd.get_semi_loss_final      (semisup_v1_sunrgbd.py:323-421)
// with their gradients.  Everything here is O(B) (one thread per frustum) except the mask cross-entropy
// (O(B*N), one CTA per frustum).  The geometric terms (8 box corners -> SUN-RGBD projection -> 2D bbox ->
// range deviation; corner distance to the ground-truth box) are differentiated by forward-mode dual
// numbers over the 7 box parameters they depend on, so the forward code is written once and the
// sub-gradient choices (min / max pick one corner, clamp, huber) follow the values exactly as autograd's do.
#pragma once
#include <initializer_list>
#include "common.cuh"

namespace t3d {

// ----------------------------------------------------------------------------- forward-mode duals
template <int NV> struct Dual {
  float v; float d[NV];
  __device__ Dual() {}
  __device__ explicit Dual(float x) : v(x) {
#pragma unroll
    for (int i = 0; i < NV; ++i) d[i] = 0.f;
  }
  __device__ static Dual var(float x, int i) { Dual r(x); r.d[i] = 1.f; return r; }
};
#define T3D_DUAL_BIN(OP, VEXPR, DEXPR)                                                  \
  template <int NV> __device__ __forceinline__ Dual<NV> OP(const Dual<NV>& a, const Dual<NV>& b) { \
    Dual<NV> r; r.v = VEXPR;                                                            \
    _Pragma("unroll") for (int i = 0; i < NV; ++i) r.d[i] = DEXPR;                      \
    return r; }
T3D_DUAL_BIN(operator+, a.v + b.v, a.d[i] + b.d[i])
T3D_DUAL_BIN(operator-, a.v - b.v, a.d[i] - b.d[i])
T3D_DUAL_BIN(operator*, a.v * b.v, a.d[i] * b.v + a.v * b.d[i])
template <int NV> __device__ __forceinline__ Dual<NV> operator/(const Dual<NV>& a, const Dual<NV>& b) {
  Dual<NV> r; const float inv = 1.0f / b.v; r.v = a.v * inv;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
template <int NV> __device__ __forceinline__ Dual<NV> operator*(const Dual<NV>& a, float s) {
  Dual<NV> r; r.v = a.v * s;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = a.d[i] * s;
  return r;
}
template <int NV> __device__ __forceinline__ Dual<NV> operator*(float s, const Dual<NV>& a) { return a * s; }
template <int NV> __device__ __forceinline__ Dual<NV> operator+(const Dual<NV>& a, float s) { Dual<NV> r = a; r.v += s; return r; }
template <int NV> __device__ __forceinline__ Dual<NV> operator-(const Dual<NV>& a, float s) { Dual<NV> r = a; r.v -= s; return r; }
template <int NV> __device__ __forceinline__ Dual<NV> operator-(const Dual<NV>& a) { return a * -1.0f; }
template <int NV> __device__ __forceinline__ Dual<NV> dsqrt(const Dual<NV>& a) {
  Dual<NV> r; r.v = sqrtf(a.v); const float g = a.v > 0.f ? 0.5f / r.v : 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = a.d[i] * g;
  return r;
}
template <int NV> __device__ __forceinline__ void dsincos(const Dual<NV>& a, Dual<NV>& s, Dual<NV>& c) {
  float sv, cv; sincosf(a.v, &sv, &cv); s.v = sv; c.v = cv;
#pragma unroll
  for (int i = 0; i < NV; ++i) { s.d[i] = cv * a.d[i]; c.d[i] = -sv * a.d[i]; }
}
// tf.losses.huber_loss / semisup_v1_sunrgbd.huber_loss on an error: 0.5*min(|e|,delta)^2 + delta*(|e| - min(|e|,delta))
template <int NV> __device__ __forceinline__ Dual<NV> dhuber(const Dual<NV>& e, float delta) {
  const float a = fabsf(e.v), q = fminf(a, delta);
  Dual<NV> r; r.v = 0.5f * q * q + delta * (a - q);
  const float g = a <= delta ? e.v : (e.v > 0.f ? delta : -delta);
#pragma unroll
  for (int i = 0; i < NV; ++i) r.d[i] = g * e.d[i];
  return r;
}
__device__ __forceinline__ float huber_f(float e, float delta, float& g) {
  const float a = fabsf(e), q = fminf(a, delta);
  g = a <= delta ? e : (e > 0.f ? delta : -delta);
  return 0.5f * q * q + delta * (a - q);
}

// ----------------------------------------------------------------------------- mask cross-entropy (O(B*N))
// out[b] = mean_n softmaxCE(logits[b,n,:], labels[b,n])   (semisup_v1_sunrgbd.py:430-431); no gradient: the
// segmentation net is not in the semi-supervised step's var_list and the mask is a non-differentiable compare.
__global__ void __launch_bounds__(256) seg_ce_kernel(const float* __restrict__ logits, const int* __restrict__ labels, int N,
                                                     float* __restrict__ out) {
  const int b = blockIdx.x;
  float s = 0.f;
  for (int n = threadIdx.x; n < N; n += 256) {
    const float2 l = *reinterpret_cast<const float2*>(logits + ((size_t)b * N + n) * 2);
    const float mx = fmaxf(l.x, l.y);
    const float lse = mx + logf(expf(l.x - mx) + expf(l.y - mx));
    s += lse - (labels[(size_t)b * N + n] ? l.y : l.x);
  }
  __shared__ float sh[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i];
    out[b] = t / (float)N;
  }
}

// soft_mask[b,n] = softmax(logits[b,n,:])[1]   (end_points['soft_mask'], semisup_v1_sunrgbd.py:103)
__global__ void __launch_bounds__(256) soft_mask_kernel(const float* __restrict__ logits, size_t n, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 l = *reinterpret_cast<const float2*>(logits + 2 * i);
  const float mx = fmaxf(l.x, l.y), e0 = expf(l.x - mx), e1 = expf(l.y - mx);
  out[i] = e1 / (e0 + e1);
}
// Gradient of the model-A loss (train_semisup.py) w.r.t. the mask logits: the cross-entropy term
// w[b] * mean_n CE(logits[b,n,:], labels[b,n]) and, when gmask != null, the surface loss through
// soft_mask = softmax(logits)[1] (gmask = d total / d soft_mask): d p1 / d l1 = p0 p1 = - d p1 / d l0.
__global__ void __launch_bounds__(256) seg_ce_bwd_kernel(const float* __restrict__ logits, const int* __restrict__ labels,
                                                         const float* __restrict__ w, const float* __restrict__ gmask, int B, int N,
                                                         float* __restrict__ dlogits) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * N) return;
  const int b = (int)(i / N);
  const float2 l = *reinterpret_cast<const float2*>(logits + 2 * i);
  const float mx = fmaxf(l.x, l.y), e0 = expf(l.x - mx), e1 = expf(l.y - mx), z = e0 + e1;
  const float p0 = e0 / z, p1 = e1 / z;
  const float wb = w[b] / (float)N;
  const int lab = labels[i];
  float g0 = wb * (p0 - (lab == 0 ? 1.f : 0.f)), g1 = wb * (p1 - (lab != 0 ? 1.f : 0.f));
  if (gmask) { const float t = gmask[i] * p0 * p1; g0 -= t; g1 += t; }
  *reinterpret_cast<float2*>(dlogits + 2 * i) = make_float2(g0, g1);
}
// out[b, c] = sum_n x[b, n, c]   (any C; the gradient of a per-frustum term broadcast over the N points: conv6's global half)
__global__ void __launch_bounds__(256) group_colsum_kernel(const float* __restrict__ x, int N, int C, int rows_per_chunk,
                                                           float* __restrict__ out) {
  const int c = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
  if (c >= C) return;
  const int n0 = blockIdx.z * rows_per_chunk, n1 = min(N, n0 + rows_per_chunk);
  const float* p = x + ((size_t)b * N + n0) * C + c;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  int n = n0;
  for (; n + 4 <= n1; n += 4, p += (size_t)4 * C) { s[0] += p[0]; s[1] += p[C]; s[2] += p[2 * (size_t)C]; s[3] += p[3 * (size_t)C]; }
  for (; n < n1; ++n, p += C) s[0] += p[0];
  atomicAdd(out + (size_t)b * C + c, (s[0] + s[1]) + (s[2] + s[3]));
}

// ----------------------------------------------------------------------------- per-class batch statistics of dims
// cls_sum[c,:] = sum of dims_reg over the samples of class c, cls_cnt[c] = their number (zero-initialised, atomics)
__global__ void class_dims_stats_kernel(const float* __restrict__ dims_reg, const float* __restrict__ one_hot, int B, int NC,
                                        float* __restrict__ cls_sum, float* __restrict__ cls_cnt) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int cid = 0; float best = one_hot[(size_t)b * NC];
  for (int c = 1; c < NC; ++c) { const float v = one_hot[(size_t)b * NC + c]; if (v > best) { best = v; cid = c; } }
  for (int k = 0; k < 3; ++k) atomicAdd(cls_sum + cid * 3 + k, dims_reg[b * 3 + k]);
  atomicAdd(cls_cnt + cid, 1.0f);
}

// ----------------------------------------------------------------------------- the semi-supervised loss
struct SemiLossArgs {
  const float* out;            // [B, 3+2NH+4NS] F_output (class_dependent/box_refine head)
  const float* stage1_center;  // [B,3]
  const float* mask_losses;    // [B] mean_N CE (seg_ce_kernel), or null
  const float* one_hot;        // [B,NC]
  const float* y_center; const int* y_orient_cls; const float* y_orient_reg; const int* y_dims_cls; const float* y_dims_reg;
  const float* Rtilt; const float* K; const float* rot_frust; const float* box2D; const float* img_dim;   // [B,9],[B,9],[B],[B,4],[B,2]
  const int* is_data_2D;       // [B]
  const float* fit_logits;     // [B,2] BoxPC fit logits of F_pred_box_reg, or null
  const float* mean_size;      // [NS,3]
  const float* cls_sum; const float* cls_cnt;    // class_dims_stats_kernel output (intra-class variance), or null
  const float* reg_in;         // [B,7] optional: evaluate the weak losses on this regression-format box instead of the one parsed from `out`
  int B, NH, NS, NC;
  unsigned icv_train_mask;     // bit c: class c is in intraclsdims_train_classes
  float w_ce, box_mult, w_center, w_ocls, w_dcls, w_oreg, w_dreg, w_tnet, w_corner;     // STRONG_*
  float weak_mult, w_icv, w_reproj, w_fit;
  int reproj_only_2d, fit_only_2d, use_softmax_proj;
  float softmax_scale, dilate;
  int clip_lower_b, clip_pred_box, reproj_mse, icv_mse;
  int train_box_mask;          // bit0 centre, bit1 dims, bit2 orient receive the reprojection gradient
  float inv_n3d;               // 1 / (sum(1 - is_data_2D) + 1e-3)
  // outputs
  float* dF;                   // [B,W]  d total / d F_output   (direct terms; box_reg_backward_kernel adds the g_reg chain)
  float* ds1;                  // [B,3]  d total / d stage1_center (direct terms)
  float* g_reg;                // [B,7]  d total / d (center_reg, dims_reg, orient_reg) from reprojection + intra-class variance
  float* dfit;                 // [B,2]  d total / d fit_logits
  float* per_sample;           // [B,6]  box_losses, corner, reproj, fit_losses, center dist, unused   (may be null)
  float* total;                // [8]    total, mask_loss, box_loss, intraclass_var, weak_loss, fit_loss, reproj mean, -- (zero-init)
};

constexpr int kMaxNH = 16, kMaxNS = 16;

template <class T> struct V3 { T x, y, z; };

// model_util.get_box3d_corners_helper (model_util.py:94-119): corner i of the box (centre c, heading h, size (l,w,h))
template <class T> __device__ __forceinline__ V3<T> helper_corner(int i, const V3<T>& c, const T& sh, const T& ch, const V3<T>& size) {
  const float sx = (i & 2) ? -0.5f : 0.5f;                                  // l,l,-l,-l,l,l,-l,-l
  const float sy = (i & 4) ? -0.5f : 0.5f;                                  // h x4, -h x4
  const float sz = ((i & 3) == 0 || (i & 3) == 3) ? 0.5f : -0.5f;           // w,-w,-w,w,...
  const T xc = size.x * sx, yc = size.z * sy, zc = size.y * sz;
  V3<T> r;
  r.x = ch * xc + sh * zc + c.x;
  r.y = yc + c.y;
  r.z = ch * zc - sh * xc + c.z;
  return r;
}

__device__ __forceinline__ float range_dev_grad(float val, float lower, float upper, int mse, float& g) {
  // weak_losses.loss_for_deviation_from_range (weak_losses.py:15-36): value and d/dval
  float loss = 0.f; g = 0.f;
  if (val < lower) { float gg; const float e = val - lower; loss += mse ? e * e : huber_f(e, 1.0f, gg); g += mse ? 2.f * e : gg; }
  if (val > upper) { float gg; const float e = val - upper; loss += mse ? e * e : huber_f(e, 1.0f, gg); g += mse ? 2.f * e : gg; }
  return loss;
}

__global__ void __launch_bounds__(64) semi_loss_kernel(const SemiLossArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (b < a.B) {
    const int NH = a.NH, NS = a.NS, W = 3 + 2 * NH + 4 * NS;
    const float* o = a.out + (size_t)b * W;
    const int base_hr = 3 + NH, base_ss = 3 + 2 * NH, base_sr = 3 + 2 * NH + NS;
    const float PI = 3.14159265358979323846f;
    const float hscale = PI / (float)NH, binw = 2.0f * PI / (float)NH;
    const int is2d = a.is_data_2D[b] != 0;
    const float ws = is2d ? 0.f : a.inv_n3d;
    const float invB = 1.0f / (float)a.B;
    float* dF = a.dF + (size_t)b * W;
    for (int k = 0; k < W; ++k) dF[k] = 0.f;
    float ds1[3] = {0, 0, 0}, dFc[3] = {0, 0, 0}, greg[7] = {0, 0, 0, 0, 0, 0, 0};
    const float s1[3] = {a.stage1_center[b * 3], a.stage1_center[b * 3 + 1], a.stage1_center[b * 3 + 2]};
    const float Fc[3] = {o[0] + s1[0], o[1] + s1[1], o[2] + s1[2]};
    const float yc[3] = {a.y_center[b * 3], a.y_center[b * 3 + 1], a.y_center[b * 3 + 2]};
    const int gt_h = a.y_orient_cls[b], gt_s = a.y_dims_cls[b];
    const float y_oreg = a.y_orient_reg[b];
    const float ydr[3] = {a.y_dims_reg[b * 3], a.y_dims_reg[b * 3 + 1], a.y_dims_reg[b * 3 + 2]};
    const float msg[3] = {a.mean_size[gt_s * 3], a.mean_size[gt_s * 3 + 1], a.mean_size[gt_s * 3 + 2]};

    // ---------------- strong loss (per sample; weights ws = (1 - is_2D) / (sum(1 - is_2D) + 1e-3))
    // centre: huber(||y - F_center||, 2);  stage-1 centre: huber(||y - stage1_center||, 1)
    float l_center, l_s1;
    {
      float e[3] = {Fc[0] - yc[0], Fc[1] - yc[1], Fc[2] - yc[2]};
      const float dist = sqrtf(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
      float g; l_center = huber_f(dist, 2.0f, g);
      const float w = ws * a.box_mult * a.w_center * (dist > 0.f ? g / dist : 0.f);
      for (int k = 0; k < 3; ++k) dFc[k] += w * e[k];
      float e1[3] = {s1[0] - yc[0], s1[1] - yc[1], s1[2] - yc[2]};
      const float d1 = sqrtf(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
      l_s1 = huber_f(d1, 1.0f, g);
      const float w1 = ws * a.box_mult * a.w_tnet * (d1 > 0.f ? g / d1 : 0.f);
      for (int k = 0; k < 3; ++k) ds1[k] += w1 * e1[k];
    }
    // heading / size classification: sparse softmax cross-entropy; argmax (first max) for the reg-format box
    float l_hcls, l_scls; int jstar = 0, istar = 0;
    {
      float mx = o[3];
      for (int j = 1; j < NH; ++j) if (o[3 + j] > mx) { mx = o[3 + j]; jstar = j; }
      float z = 0.f;
      for (int j = 0; j < NH; ++j) z += expf(o[3 + j] - mx);
      l_hcls = logf(z) + mx - o[3 + gt_h];
      const float w = ws * a.box_mult * a.w_ocls;
      for (int j = 0; j < NH; ++j) dF[3 + j] += w * (expf(o[3 + j] - mx) / z - (j == gt_h ? 1.f : 0.f));
      mx = o[base_ss];
      for (int i = 1; i < NS; ++i) if (o[base_ss + i] > mx) { mx = o[base_ss + i]; istar = i; }
      z = 0.f;
      for (int i = 0; i < NS; ++i) z += expf(o[base_ss + i] - mx);
      l_scls = logf(z) + mx - o[base_ss + gt_s];
      const float w2 = ws * a.box_mult * a.w_dcls;
      for (int i = 0; i < NS; ++i) dF[base_ss + i] += w2 * (expf(o[base_ss + i] - mx) / z - (i == gt_s ? 1.f : 0.f));
    }
    // residuals at the ground-truth bins
    float l_hreg, l_sreg;
    {
      float g;
      l_hreg = huber_f(o[base_hr + gt_h] - y_oreg / hscale, 1.0f, g);
      dF[base_hr + gt_h] += ws * a.box_mult * a.w_oreg * g;
      float e[3]; float n2 = 0.f;
      for (int k = 0; k < 3; ++k) { e[k] = ydr[k] / msg[k] - o[base_sr + gt_s * 3 + k]; n2 += e[k] * e[k]; }
      const float dist = sqrtf(n2);
      l_sreg = huber_f(dist, 1.0f, g);
      const float w = ws * a.box_mult * a.w_dreg * (dist > 0.f ? g / dist : 0.f);
      for (int k = 0; k < 3; ++k) dF[base_sr + gt_s * 3 + k] -= w * e[k];
    }
    // corner loss: the GT-(heading bin, size cluster) candidate of get_box3d_corners_sunrgbd (size residual added twice,
    // model_util.py:158-159) against the GT box and its pi-flipped copy; variables: F_center(3), heading residual, size residual(3)
    float l_corner;
    {
      typedef Dual<7> D;
      V3<D> c; c.x = D::var(Fc[0], 0); c.y = D::var(Fc[1], 1); c.z = D::var(Fc[2], 2);
      const float hr = o[base_hr + gt_h] * hscale;
      D heading = D::var(hr + binw * (float)gt_h, 3);
      V3<D> size;
      size.x = D::var(msg[0] + 2.f * (o[base_sr + gt_s * 3 + 0] * msg[0]), 4); size.x.d[4] = 2.f;
      size.y = D::var(msg[1] + 2.f * (o[base_sr + gt_s * 3 + 1] * msg[1]), 5); size.y.d[5] = 2.f;
      size.z = D::var(msg[2] + 2.f * (o[base_sr + gt_s * 3 + 2] * msg[2]), 6); size.z.d[6] = 2.f;
      D sh, ch; dsincos(heading, sh, ch);
      const float hl = y_oreg + binw * (float)gt_h;
      float shl, chl; sincosf(hl, &shl, &chl);
      float shf, chf; sincosf(hl + PI, &shf, &chf);
      V3<float> cg; cg.x = yc[0]; cg.y = yc[1]; cg.z = yc[2];
      V3<float> sg; sg.x = msg[0] + ydr[0]; sg.y = msg[1] + ydr[1]; sg.z = msg[2] + ydr[2];
      D sum(0.f);
      for (int i = 0; i < 8; ++i) {
        const V3<D> p = helper_corner<D>(i, c, sh, ch, size);
        const V3<float> g1 = helper_corner<float>(i, cg, shl, chl, sg);
        const V3<float> g2 = helper_corner<float>(i, cg, shf, chf, sg);
        const D ax = p.x - g1.x, ay = p.y - g1.y, az = p.z - g1.z;
        const D bx = p.x - g2.x, by = p.y - g2.y, bz = p.z - g2.z;
        const D d1 = dsqrt(ax * ax + ay * ay + az * az), d2 = dsqrt(bx * bx + by * by + bz * bz);
        sum = sum + dhuber(d1.v <= d2.v ? d1 : d2, 1.0f);
      }
      sum = sum * 0.125f;
      l_corner = sum.v;
      const float w = ws * a.w_corner;
      for (int k = 0; k < 3; ++k) dFc[k] += w * sum.d[k];
      dF[base_hr + gt_h] += w * sum.d[3] * hscale;
      for (int k = 0; k < 3; ++k) dF[base_sr + gt_s * 3 + k] += w * sum.d[4 + k] * msg[k];
    }
    const float box_loss = a.box_mult * (a.w_center * l_center + a.w_ocls * l_hcls + a.w_dcls * l_scls + a.w_oreg * l_hreg +
                                         a.w_dreg * l_sreg + a.w_tnet * l_s1) + a.w_corner * l_corner;
    const float mask_l = a.mask_losses ? a.w_ce * a.mask_losses[b] : 0.f;
    acc[1] = ws * mask_l; acc[2] = ws * box_loss;

    // ---------------- box in regression format (tf_convert_box_params_from_anchor_to_reg_format, tf_util.py:1001-1041)
    float dims[3]; bool dims_live[3];
    for (int k = 0; k < 3; ++k) {
      const float m = a.mean_size[istar * 3 + k];
      const float v = m + o[base_sr + istar * 3 + k] * m;
      dims_live[k] = v >= 1e-5f; dims[k] = fmaxf(v, 1e-5f);
    }
    float orient = binw * (float)jstar + o[base_hr + jstar] * hscale;
    float rc[3] = {Fc[0], Fc[1], Fc[2]};
    if (a.reg_in) {
      const float* r = a.reg_in + (size_t)b * 7;
      for (int k = 0; k < 3; ++k) { rc[k] = r[k]; dims[k] = r[3 + k]; dims_live[k] = true; }
      orient = r[6];
    }

    // ---------------- intra-class variance of dims_reg (weak_losses.py:267-291); group means are stop-gradient
    int cid = 0;
    { float best = a.one_hot[(size_t)b * a.NC]; for (int c = 1; c < a.NC; ++c) { const float v = a.one_hot[(size_t)b * a.NC + c]; if (v > best) { best = v; cid = c; } } }
    if (a.w_icv != 0.f && a.cls_sum && ((a.icv_train_mask >> cid) & 1u)) {
      const float cnt = a.cls_cnt[cid];
      const float T = (float)__popc(a.icv_train_mask & ((1u << a.NC) - 1u));
      const float norm = 1.0f / (3.0f * cnt * T);
      for (int k = 0; k < 3; ++k) {
        const float e = dims[k] - a.cls_sum[cid * 3 + k] / cnt;
        float g; const float l = a.icv_mse ? e * e : huber_f(e, 1.0f, g);
        if (a.icv_mse) g = 2.f * e;
        acc[3] += l * norm;
        greg[3 + k] += a.weak_mult * a.w_icv * g * norm;
      }
    }

    // ---------------- relaxed reprojection loss (weak_losses.py:69-238); variables: centre(3), dims(3), orient
    float l_reproj = 0.f;
    if (a.w_reproj != 0.f) {
      typedef Dual<7> D;
      D cx = D::var(rc[0], 0), cy = D::var(rc[1], 1), cz = D::var(rc[2], 2);
      D dl = D::var(dims[0], 3), dw = D::var(dims[1], 4), dh = D::var(dims[2], 5), th = D::var(orient, 6);
      if (!(a.train_box_mask & 1)) { cx.d[0] = 0.f; cy.d[1] = 0.f; cz.d[2] = 0.f; }
      if (!(a.train_box_mask & 2)) { dl.d[3] = 0.f; dw.d[4] = 0.f; dh.d[5] = 0.f; }
      if (!(a.train_box_mask & 4)) th.d[6] = 0.f;
      // tf_rot_box_params_multi (tf_util.py:1045-1073) by +rot_frust
      const float ang = a.rot_frust[b];
      float sa, ca; sincosf(ang, &sa, &ca);
      const D ncx = cx * ca + cz * sa, ncy = cy, ncz = cz * ca - cx * sa;
      const D th2 = th + ang;
      D s, c; dsincos(-th2, s, c);
      const float* R = a.Rtilt + (size_t)b * 9;
      const float* Km = a.K + (size_t)b * 9;
      D px[8], py[8];
      for (int i = 0; i < 8; ++i) {
        // tf_create_3D_box_by_vertices_multi (tf_util.py:841-891): upright depth corners, rotz(-theta), flip to upright camera
        const float sx = (i == 0 || i == 3 || i == 4 || i == 7) ? -0.5f : 0.5f;
        const float sy = ((i & 3) < 2) ? 0.5f : -0.5f;
        const float sz = (i < 4) ? 0.5f : -0.5f;
        const D xc = dl * sx, yc_ = dw * sy, zc = dh * sz;
        const D xd = c * xc - s * yc_, yd = s * xc + c * yc_;
        const D camx = xd + ncx, camy = -zc + ncy, camz = yd + ncz;
        // project (tf_util.py:798-838): camera -> depth (x, z, -y); Rtilt^T; depth -> camera (x, -z, y); K; perspective divide
        const D d0 = camx, d1 = camz, d2 = -camy;
        const D q0 = d0 * R[0] + d1 * R[3] + d2 * R[6];
        const D q1 = d0 * R[1] + d1 * R[4] + d2 * R[7];
        const D q2 = d0 * R[2] + d1 * R[5] + d2 * R[8];
        const D c0 = q0, c1 = -q2, c2 = q1;
        const D u = c0 * Km[0] + c1 * Km[1] + c2 * Km[2];
        const D v = c0 * Km[3] + c1 * Km[4] + c2 * Km[5];
        const D w = c0 * Km[6] + c1 * Km[7] + c2 * Km[8];
        px[i] = u / w; py[i] = v / w;
      }
      D pb[4];
      {
        int il = 0, ir = 0, it = 0, ib = 0;
        for (int i = 1; i < 8; ++i) {
          if (px[i].v < px[il].v) il = i;
          if (px[i].v > px[ir].v) ir = i;
          if (py[i].v < py[it].v) it = i;
          if (py[i].v > py[ib].v) ib = i;
        }
        if (!a.use_softmax_proj) { pb[0] = px[il]; pb[1] = py[it]; pb[2] = px[ir]; pb[3] = py[ib]; }
        else {
          // tf_get_2D_softmax_bbox_of_points (tf_util.py:379-413): softmax weights and width/height are stop-gradient
          const float lb = px[il].v, rb = px[ir].v, tb = py[it].v, bb = py[ib].v;
          const float wdt = fabsf(rb - lb), hgt = fabsf(bb - tb), sc = a.softmax_scale;
          for (int side = 0; side < 4; ++side) {
            float z[8], mx = -3.0e38f, zs = 0.f;
            for (int i = 0; i < 8; ++i) {
              z[i] = side == 0 ? (rb - px[i].v) / wdt : side == 1 ? (bb - py[i].v) / hgt : side == 2 ? (px[i].v - lb) / wdt : (py[i].v - tb) / hgt;
              z[i] *= sc; mx = fmaxf(mx, z[i]);
            }
            for (int i = 0; i < 8; ++i) { z[i] = expf(z[i] - mx); zs += z[i]; }
            D acc2(0.f);
            for (int i = 0; i < 8; ++i) acc2 = acc2 + ((side & 1) ? py[i] : px[i]) * (z[i] / zs);
            pb[side] = acc2;
          }
        }
      }
      // bounds from the 2D box: small = clip(box2D), big = dilate(box2D), big_clip = clip(big)   (tf_util.py:486-540)
      const float* bx = a.box2D + (size_t)b * 4;
      const float rows = a.img_dim[b * 2], cols = a.img_dim[b * 2 + 1];
      const float small[4] = {fmaxf(0.f, bx[0]), fmaxf(0.f, bx[1]), fminf(cols, bx[2]), fminf(rows, bx[3])};
      const float ccx = (bx[0] + bx[2]) * 0.5f, ccy = (bx[1] + bx[3]) * 0.5f;
      const float nw = a.dilate * (bx[2] - bx[0]), nh = a.dilate * (bx[1] - bx[3]);
      const float big[4] = {ccx - nw * 0.5f, ccy + nh * 0.5f, ccx + nw * 0.5f, ccy - nh * 0.5f};
      const float bigc[4] = {fmaxf(0.f, big[0]), fmaxf(0.f, big[1]), fminf(cols, big[2]), fminf(rows, big[3])};
      float gside[4] = {0, 0, 0, 0};
      if (a.clip_pred_box) {
        // predicted box clipped to the image as well; gradient passes only where the clip is inactive
        for (int k = 0; k < 4; ++k) {
          float v = pb[k].v; bool live = true;
          if (k < 2) { if (v < 0.f) { v = 0.f; live = false; } }
          else { const float lim = k == 2 ? cols : rows; if (v > lim) { v = lim; live = false; } }
          const float lo = k < 2 ? bigc[k] : small[k], hi = k < 2 ? small[k] : bigc[k];
          float g; l_reproj += range_dev_grad(v, lo, hi, a.reproj_mse, g);
          gside[k] = live ? g : 0.f;
        }
      } else if (a.clip_lower_b) {
        for (int k = 0; k < 4; ++k) {
          const float nc = big[k] == bigc[k] ? 1.f : 0.f;
          const float lo = k < 2 ? bigc[k] : small[k], hi = k < 2 ? small[k] : bigc[k];
          float g; l_reproj += nc * range_dev_grad(pb[k].v, lo, hi, a.reproj_mse, g);
          gside[k] = nc * g;
        }
        if (l_reproj > 1000.f) { l_reproj = 1000.f; for (int k = 0; k < 4; ++k) gside[k] = 0.f; }
      } else {
        for (int k = 0; k < 4; ++k) {
          const float nc = big[k] == bigc[k] ? 1.f : 0.f;
          const float v = pb[k].v;
          float gi, go, li, lo2;
          { const float e = v - small[k]; li = a.reproj_mse ? e * e : huber_f(e, 1.0f, gi); if (a.reproj_mse) gi = 2.f * e; }
          { const float e = v - bigc[k]; lo2 = a.reproj_mse ? e * e : huber_f(e, 1.0f, go); if (a.reproj_mse) go = 2.f * e; }
          if (k < 2) { if (v < small[k]) { li = 0.f; gi = 0.f; } if (v > bigc[k]) { lo2 = 0.f; go = 0.f; } }
          else { if (v > small[k]) { li = 0.f; gi = 0.f; } if (v < bigc[k]) { lo2 = 0.f; go = 0.f; } }
          float l = li + lo2 * nc, g = gi + go * nc;
          if (l > 1000.f) { l = 1000.f; g = 0.f; }
          l_reproj += l; gside[k] = g;
        }
      }
      const float wr = a.weak_mult * a.w_reproj * ((a.reproj_only_2d && !is2d) ? 0.f : 1.f) * invB;
      for (int v = 0; v < 7; ++v) {
        float t = 0.f;
        for (int k = 0; k < 4; ++k) t += gside[k] * pb[k].d[v];
        greg[v] += wr * t;
      }
      acc[6] = l_reproj * ((a.reproj_only_2d && !is2d) ? 0.f : 1.f) * invB;
    }

    // ---------------- BoxPC fit loss: -log(0.01 + softmax(fit_logits)[1])   (semisup_v1_sunrgbd.py:399-411)
    float l_fit = 0.f;
    if (a.w_fit != 0.f && a.fit_logits) {
      const float l0 = a.fit_logits[b * 2], l1 = a.fit_logits[b * 2 + 1], mx = fmaxf(l0, l1);
      const float e0 = expf(l0 - mx), e1 = expf(l1 - mx), p = e1 / (e0 + e1);
      l_fit = -logf(0.01f + p);
      const float wf = a.w_fit * ((a.fit_only_2d && !is2d) ? 0.f : 1.f) * invB;
      const float dp = -wf / (0.01f + p);
      if (a.dfit) { a.dfit[b * 2] = -dp * p * (1.f - p); a.dfit[b * 2 + 1] = dp * p * (1.f - p); }
      acc[5] = l_fit * ((a.fit_only_2d && !is2d) ? 0.f : 1.f) * invB;
    } else if (a.dfit) { a.dfit[b * 2] = 0.f; a.dfit[b * 2 + 1] = 0.f; }

    // F_center = F_output[0:3] + stage1_center: both receive its gradient
    for (int k = 0; k < 3; ++k) { dF[k] += dFc[k]; ds1[k] += dFc[k]; }
    for (int k = 0; k < 3; ++k) a.ds1[b * 3 + k] = ds1[k];
    for (int k = 0; k < 7; ++k) a.g_reg[b * 7 + k] = (k >= 3 && k < 6 && !dims_live[k - 3]) ? 0.f : greg[k];
    if (a.per_sample) {
      float* ps = a.per_sample + (size_t)b * 6;
      ps[0] = box_loss; ps[1] = l_corner; ps[2] = l_reproj; ps[3] = l_fit; ps[4] = l_center; ps[5] = mask_l;
    }
  }
  // weak_loss = mean_B(w_icv * icv + w_reproj * reproj [* is_2D]); total = strong + mult * weak + w_fit * fit
  acc[4] = a.w_icv * acc[3] + a.w_reproj * acc[6];
  acc[0] = acc[1] + acc[2] + a.weak_mult * acc[4] + a.w_fit * acc[5];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
    if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(a.total + k, v);
  }
}

// dF += chain of g_reg through tf_convert_box_params_from_anchor_to_reg_format (argmax-selected residuals) and the parse
// scalings; ds1 += g_reg[0:3].  g_reg = (d centre(3), d dims(3), d orient).
__global__ void box_reg_backward_kernel(const float* __restrict__ out, const float* __restrict__ g_reg, const float* __restrict__ mean_size,
                                        int B, int NH, int NS, float* __restrict__ dF, float* __restrict__ ds1) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int W = 3 + 2 * NH + 4 * NS, base_hr = 3 + NH, base_ss = 3 + 2 * NH, base_sr = base_ss + NS;
  const float* o = out + (size_t)b * W;
  const float* g = g_reg + (size_t)b * 7;
  int jstar = 0, istar = 0;
  for (int j = 1; j < NH; ++j) if (o[3 + j] > o[3 + jstar]) jstar = j;
  for (int i = 1; i < NS; ++i) if (o[base_ss + i] > o[base_ss + istar]) istar = i;
  float* d = dF + (size_t)b * W;
  for (int k = 0; k < 3; ++k) { d[k] += g[k]; if (ds1) ds1[b * 3 + k] += g[k]; }
  for (int k = 0; k < 3; ++k) {
    const float m = mean_size[istar * 3 + k];
    if (m + o[base_sr + istar * 3 + k] * m >= 1e-5f) d[base_sr + istar * 3 + k] += g[3 + k] * m;
  }
  d[base_hr + jstar] += g[6] * (3.14159265358979323846f / (float)NH);
}

// ----------------------------------------------------------------------------- BoxPC representation backward
// d (centre, dims, orient) of tf_get_box_pc_representation (tf_util.py:764-795) from the gradient of its 6 plane-distance
// channels; one CTA per frustum.  g6: [B*N, 6].  g_box: [B,7], accumulated (+=).
__global__ void __launch_bounds__(256) boxpc_features_bwd_kernel(const float* __restrict__ pc, int N, int C, const float* __restrict__ center,
                                                                 const float* __restrict__ orient, const float* __restrict__ g6,
                                                                 float* __restrict__ g_box) {
  const int b = blockIdx.x;
  float st, ct; sincosf(orient[b], &st, &ct);
  const float cx = center[b * 3], cz = center[b * 3 + 2];
  float acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int n = threadIdx.x; n < N; n += 256) {
    const float* q = pc + ((size_t)b * N + n) * C;
    const float* g = g6 + ((size_t)b * N + n) * 6;
    const float dx = q[0] - cx, dz = q[2] - cz;
    const float xr = ct * dx - st * dz, zr = st * dx + ct * dz;
    const float ga = g[1] - g[0], gb = g[3] - g[2], gc = g[5] - g[4];      // coefficients of xr, dy, zr
    acc[0] -= ga * ct + gc * st;
    acc[1] -= gb;
    acc[2] += ga * st - gc * ct;
    acc[3] += 0.5f * (g[0] + g[1]);          // l
    acc[4] += 0.5f * (g[4] + g[5]);          // w
    acc[5] += 0.5f * (g[2] + g[3]);          // h
    acc[6] += gc * xr - ga * zr;             // d xr / d theta = -zr, d zr / d theta = xr
  }
  __shared__ float sh[8][7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
    g_box[b * 7 + threadIdx.x] += t;
  }
}

// ----------------------------------------------------------------------------- small element-wise helpers
// dOut *= act'(out)   (eval-mode layer backward; act: 1 relu, 2 leaky_relu(0.2), 3 tanh)
__global__ void act_bwd_kernel(float* __restrict__ dout, const float* __restrict__ out, size_t n, int act) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float y = out[i];
  float g = 1.f;
  if (act == 1) g = y > 0.f ? 1.f : 0.f;
  else if (act == 2) g = y > 0.f ? 1.f : 0.2f;
  else if (act == 3) g = 1.f - y * y;
  dout[i] *= g;
}
// out[r, c] = x[r, c] * rowmask[r]    (net * mask before the max-pool, semisup_models.py:184-185,240-241, and its backward)
__global__ void rowmask_mul_kernel(const float* __restrict__ x, const float* __restrict__ rowmask, float* __restrict__ out, size_t total, int C) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) out[i] = x[i] * rowmask[i / C];
}

// ----------------------------------------------------------------------------- 128-bit variants of the streaming kernels
// The training step is half HBM-bound element-wise passes (ncu launch list, profiles/r01_launches_bench_cfg5_tc.csv); the
// scalar kernels above ran at ~2.4 TB/s (one element per thread, a 64-bit modulo per element).  These move float4s, four
// per thread with all loads issued before the first use, and take the channel from a 32-bit modulo on the float4 index.
// Same arithmetic expressions as the scalar kernels (bit-identical results); used when C % 4 == 0 and the pointers are
// 16-byte aligned.
constexpr int kEw4 = 4;      // float4s per thread
#define T3D_EW4_LOOP(q) for (int u = 0; u < kEw4; ++u) if (const unsigned q = blockIdx.x * (256u * kEw4) + u * 256u + threadIdx.x; q < n4)

// per-channel parameters live at arbitrary offsets of the flat parameter arena: scalar loads (L1 hits), no alignment demand
__device__ __forceinline__ float4 ldg4(const float* __restrict__ p, unsigned c4) {
  const float* q = p + 4u * c4;
  return make_float4(__ldg(q), __ldg(q + 1), __ldg(q + 2), __ldg(q + 3));
}
// Per-channel parameters are staged once per block in shared memory (as float4s, 16 scalar loads per data float4 otherwise)
// and a block streams kBnIter x 1024 float4s (64 KB) to amortise the staging.
constexpr int kBnIter = 4;
__global__ void __launch_bounds__(256) bn_apply4_kernel(const float4* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, float4* __restrict__ out,
                                                       unsigned n4, unsigned C4, int act) {
  extern __shared__ float4 bn_sp[];                  // [4][C4]: mean, rstd, gamma, beta
  for (unsigned i = threadIdx.x; i < C4; i += 256) {
    bn_sp[i] = ldg4(mean, i); bn_sp[C4 + i] = ldg4(rstd, i); bn_sp[2 * C4 + i] = ldg4(gamma, i); bn_sp[3 * C4 + i] = ldg4(beta, i);
  }
  __syncthreads();
#pragma unroll 1
  for (int it = 0; it < kBnIter; ++it) {
    const unsigned base = (blockIdx.x * kBnIter + it) * (256u * kEw4) + threadIdx.x;
    float4 v[kEw4];
#pragma unroll
    for (int u = 0; u < kEw4; ++u) if (base + u * 256u < n4) v[u] = y[base + u * 256u];
#pragma unroll
    for (int u = 0; u < kEw4; ++u) {
      const unsigned q = base + u * 256u;
      if (q >= n4) continue;
      const unsigned c = q % C4;
      const float4 m = bn_sp[c], r = bn_sp[C4 + c], g = bn_sp[2 * C4 + c], b = bn_sp[3 * C4 + c];
      float4 o;
      o.x = act_apply(g.x * (v[u].x - m.x) * r.x + b.x, act);
      o.y = act_apply(g.y * (v[u].y - m.y) * r.y + b.y, act);
      o.z = act_apply(g.z * (v[u].z - m.z) * r.z + b.z, act);
      o.w = act_apply(g.w * (v[u].w - m.w) * r.w + b.w, act);
      out[q] = o;
    }
  }
}

// has_out == 2: lazy BN layer, o carries a_scale * y + a_shift (the forward ReLU argument) instead of the stored output
__device__ __forceinline__ float bn_bwd_one(float dy, float o, int has_out, float y, float m, float r, float g, float s1, float s2, float inv, int act) {
  if (has_out == 1) dy *= act_grad_from_out(o, act);
  else if (has_out == 2) dy = o > 0.0f ? dy : 0.0f;
  const float xh = (y - m) * r;
  return g * r * (dy - s1 * inv - xh * s2 * inv);
}
__global__ void __launch_bounds__(256) bn_backward4_kernel(float4* __restrict__ dOut, const float4* __restrict__ out, const float4* __restrict__ y,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          const float* __restrict__ gamma, const float* __restrict__ s1,
                                                          const float* __restrict__ s2, unsigned n4, unsigned C4, int M, int act,
                                                          const float* __restrict__ a_scale = nullptr,
                                                          const float* __restrict__ a_shift = nullptr) {
  extern __shared__ float4 bn_sp[];                  // [5 or 7][C4]: mean, rstd, gamma, s1, s2 (, a_scale, a_shift)
  const int has_out = out != nullptr ? 1 : (a_scale != nullptr ? 2 : 0);
  for (unsigned i = threadIdx.x; i < C4; i += 256) {
    bn_sp[i] = ldg4(mean, i); bn_sp[C4 + i] = ldg4(rstd, i); bn_sp[2 * C4 + i] = ldg4(gamma, i);
    bn_sp[3 * C4 + i] = ldg4(s1, i); bn_sp[4 * C4 + i] = ldg4(s2, i);
    if (has_out == 2) { bn_sp[5 * C4 + i] = ldg4(a_scale, i); bn_sp[6 * C4 + i] = ldg4(a_shift, i); }
  }
  __syncthreads();
  const float inv = 1.0f / (float)M;
#pragma unroll 1
  for (int it = 0; it < kBnIter; ++it) {
    const unsigned base = (blockIdx.x * kBnIter + it) * (256u * kEw4) + threadIdx.x;
    float4 d[kEw4], o[kEw4], yy[kEw4];
#pragma unroll
    for (int u = 0; u < kEw4; ++u) {
      const unsigned q = base + u * 256u;
      if (q < n4) { d[u] = dOut[q]; yy[u] = y[q]; o[u] = has_out == 1 ? out[q] : make_float4(0.f, 0.f, 0.f, 0.f); }
    }
#pragma unroll
    for (int u = 0; u < kEw4; ++u) {
      const unsigned q = base + u * 256u;
      if (q >= n4) continue;
      const unsigned c = q % C4;
      const float4 m = bn_sp[c], r = bn_sp[C4 + c], g = bn_sp[2 * C4 + c], a1 = bn_sp[3 * C4 + c], a2 = bn_sp[4 * C4 + c];
      if (has_out == 2) {
        const float4 sc = bn_sp[5 * C4 + c], sh = bn_sp[6 * C4 + c];
        o[u] = make_float4(fmaf(sc.x, yy[u].x, sh.x), fmaf(sc.y, yy[u].y, sh.y), fmaf(sc.z, yy[u].z, sh.z), fmaf(sc.w, yy[u].w, sh.w));
      }
      float4 w;
      w.x = bn_bwd_one(d[u].x, o[u].x, has_out, yy[u].x, m.x, r.x, g.x, a1.x, a2.x, inv, act);
      w.y = bn_bwd_one(d[u].y, o[u].y, has_out, yy[u].y, m.y, r.y, g.y, a1.y, a2.y, inv, act);
      w.z = bn_bwd_one(d[u].z, o[u].z, has_out, yy[u].z, m.z, r.z, g.z, a1.z, a2.z, inv, act);
      w.w = bn_bwd_one(d[u].w, o[u].w, has_out, yy[u].w, m.w, r.w, g.w, a1.w, a2.w, inv, act);
      dOut[q] = w;
    }
  }
}

__global__ void __launch_bounds__(256) act_bwd4_kernel(float4* __restrict__ dout, const float4* __restrict__ out, unsigned n4, int act) {
  float4 d[kEw4], o[kEw4];
#pragma unroll
  T3D_EW4_LOOP(q) { d[u] = dout[q]; o[u] = out[q]; }
#pragma unroll
  T3D_EW4_LOOP(q) {
    float4 w = d[u];
    w.x *= act_grad_from_out(o[u].x, act); w.y *= act_grad_from_out(o[u].y, act);
    w.z *= act_grad_from_out(o[u].z, act); w.w *= act_grad_from_out(o[u].w, act);
    dout[q] = w;
  }
}

__global__ void __launch_bounds__(256) rowmask_mul4_kernel(const float4* __restrict__ x, const float* __restrict__ rowmask, float4* __restrict__ out,
                                                          unsigned n4, unsigned C4) {
  float4 v[kEw4]; float rm[kEw4];
#pragma unroll
  T3D_EW4_LOOP(q) { v[u] = x[q]; rm[u] = __ldg(rowmask + q / C4); }
#pragma unroll
  T3D_EW4_LOOP(q) out[q] = make_float4(v[u].x * rm[u], v[u].y * rm[u], v[u].z * rm[u], v[u].w * rm[u]);
}

__global__ void __launch_bounds__(256) scale_mask4_kernel(const float4* __restrict__ x, const float4* __restrict__ mask, float scale,
                                                         float4* __restrict__ out, unsigned n4) {
  float4 v[kEw4], m[kEw4];
#pragma unroll
  T3D_EW4_LOOP(q) { v[u] = x[q]; m[u] = mask[q]; }
#pragma unroll
  T3D_EW4_LOOP(q) out[q] = make_float4(v[u].x * m[u].x * scale, v[u].y * m[u].y * scale, v[u].z * m[u].z * scale, v[u].w * m[u].w * scale);
}
#undef T3D_EW4_LOOP
inline bool ew4_ok(size_t total, int C, std::initializer_list<const void*> ptrs) {
  if (C % 4 != 0 || total % 4 != 0 || total / 4 >= 0xffffffffull) return false;
  for (const void* p : ptrs) if (p && ((uintptr_t)p & 15)) return false;
  return true;
}
inline unsigned ew4_grid(size_t total) { return (unsigned)((total / 4 + 256 * kEw4 - 1) / (256 * kEw4)); }

// out[b, c] = scale * sum_n x[b, n, c]   (C <= 8; e.g. d stage1_center = -sum_n d(xyz - stage1_center), semisup_models.py:204-209)
__global__ void __launch_bounds__(256) group_sum_kernel(const float* __restrict__ x, int N, int C, float scale, float* __restrict__ out) {
  const int b = blockIdx.x;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int n = threadIdx.x; n < N; n += 256)
    for (int c = 0; c < C; ++c) acc[c] += x[((size_t)b * N + n) * C + c];
  __shared__ float sh[8][8];
  for (int c = 0; c < C; ++c) {
    float v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][c] = v;
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
    out[b * C + threadIdx.x] = scale * t;
  }
}

}  // namespace t3d
