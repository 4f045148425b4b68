// Inference post-processing: the numpy block behind test_semisup.inference
// (sunrgbd_detection/test_semisup.py:236-258) and roi_seg_box3d_dataset.from_prediction_to_label_format (:461-466) as
// two kernels, so that what leaves the device per frustum is the prediction (a byte per point + ~20 floats) instead of
// the raw fetches (2 floats per point + 67 floats) the reference copies back and reduces on the host.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace t3d {

struct InferScoreArgs {
  const float* logits;              // [B,N,2]
  const float* heading_scores;      // [B,NH]
  const float* heading_residuals;   // [B,NH]
  const float* size_scores;         // [B,NS]
  const float* size_residuals;      // [B,NS,3]
  const float* fit_prob;            // [B] or null (use_boxpc_fit_prob)
  int B, N, NH, NS;
  uint8_t* pred_seg;                // [B,N]  argmax(logits, 2): 1 iff l1 > l0 (np.argmax takes the first maximum)
  float* mask_mean_prob;            // [B]    sum(prob1 * seg) / (sum(seg) + 1)
  int* heading_cls;                 // [B]    argmax(heading_scores)
  float* heading_res;               // [B]    heading_residuals[b, heading_cls[b]]
  int* size_cls;                    // [B]
  float* size_res;                  // [B,3]
  float* scores;                    // [B]    log(mask_mean_prob + .01) + log(max softmax(heading) + .01) + log(max softmax(size) + .01) [+ log(fit + .01)]
};

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += sh[w];
  __syncthreads();
  return t;
}

// max softmax probability and first argmax of a short vector (one thread)
__device__ __forceinline__ void softmax_max(const float* x, int n, float& pmax, int& amax) {
  float m = x[0]; int a = 0;
  for (int i = 1; i < n; ++i) if (x[i] > m) { m = x[i]; a = i; }
  float z = 0.f;
  for (int i = 0; i < n; ++i) z += expf(x[i] - m);
  pmax = 1.0f / z; amax = a;
}

// one CTA (256 threads) per frustum; logits are read once (8 B / point), pred_seg written once (1 B / point)
__global__ void __launch_bounds__(256) inference_scores_kernel(const InferScoreArgs a) {
  __shared__ float sh[8];
  const int b = blockIdx.x;
  const float2* lg = reinterpret_cast<const float2*>(a.logits) + (size_t)b * a.N;
  uint8_t* seg = a.pred_seg ? a.pred_seg + (size_t)b * a.N : nullptr;
  float sp = 0.f, cnt = 0.f;
  for (int n = threadIdx.x; n < a.N; n += 256) {
    const float2 l = lg[n];
    const float m = fmaxf(l.x, l.y);
    const float e0 = expf(l.x - m), e1 = expf(l.y - m);
    const float p1 = e1 / (e0 + e1);
    const bool in = l.y > l.x;
    if (in) { sp += p1; cnt += 1.f; }
    if (seg) seg[n] = in ? 1 : 0;
  }
  sp = block_sum_256(sp, sh);
  cnt = block_sum_256(cnt, sh);
  if (threadIdx.x == 0) {
    const float mmp = sp / (cnt + 1.0f);
    float hp, zp; int hc, zc;
    softmax_max(a.heading_scores + (size_t)b * a.NH, a.NH, hp, hc);
    softmax_max(a.size_scores + (size_t)b * a.NS, a.NS, zp, zc);
    float s = logf(mmp + 0.01f) + logf(hp + 0.01f) + logf(zp + 0.01f);
    if (a.fit_prob) s += logf(a.fit_prob[b] + 0.01f);
    if (a.mask_mean_prob) a.mask_mean_prob[b] = mmp;
    a.heading_cls[b] = hc;
    a.heading_res[b] = a.heading_residuals[(size_t)b * a.NH + hc];
    a.size_cls[b] = zc;
    for (int k = 0; k < 3; ++k) a.size_res[(size_t)b * 3 + k] = a.size_residuals[((size_t)b * a.NS + zc) * 3 + k];
    a.scores[b] = s;
  }
}

// from_prediction_to_label_format for a batch: out[b] = (h, w, l, tx, ty, tz, ry)
//   (l, w, h) = mean_size[size_cls] + size_res;  ry = class2angle(heading_cls, heading_res, NH) + rot_angle, where class2angle
//   wraps angles > pi by -2 pi before the rotation is added;  (tx, tz) = the centre rotated by -rot_angle about y
//   (rotate_pc_along_y: x' = c x - s z, z' = s x + c z with c = cos(-rot), s = sin(-rot));  ty = cy + h / 2.
__global__ void prediction_to_label_kernel(const float* __restrict__ center, const int* __restrict__ heading_cls,
                                           const float* __restrict__ heading_res, const int* __restrict__ size_cls,
                                           const float* __restrict__ size_res, const float* __restrict__ rot_angle,
                                           const float* __restrict__ mean_size, int B, int NH, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int sc = size_cls[b];
  const float l = mean_size[sc * 3 + 0] + size_res[b * 3 + 0];
  const float w = mean_size[sc * 3 + 1] + size_res[b * 3 + 1];
  const float h = mean_size[sc * 3 + 2] + size_res[b * 3 + 2];
  const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
  float ang = (float)heading_cls[b] * (two_pi / (float)NH) + heading_res[b];
  if (ang > pi) ang -= two_pi;
  const float rot = rot_angle[b];
  float s, c;
  sincosf(-rot, &s, &c);
  const float x = center[b * 3 + 0], y = center[b * 3 + 1], z = center[b * 3 + 2];
  float* o = out + (size_t)b * 7;
  o[0] = h; o[1] = w; o[2] = l;
  o[3] = c * x - s * z;
  o[4] = y + h / 2.0f;
  o[5] = s * x + c * z;
  o[6] = ang + rot;
}

}  // namespace t3d
