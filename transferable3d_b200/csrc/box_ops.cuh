// Oriented 3D box IoU and the IoU-band box perturbation of the BoxPC-Fit data path.
//
// The reference imports `box_util.box3d_iou` (roi_seg_box3d_dataset.py:15, box_pc_fit_dataset.py:17) but the module is
// absent from its tree; it is the box_util.py of charlesq34/frustum-pointnets (train/box_util.py), whose published
// algorithm is restated here and in oracle/box_util.py:
//   boxes are upright (rotation about y only): bird's-eye-view rectangles (x, z of corners 3,2,1,0) are intersected by
//   Sutherland-Hodgman clipping, the polygon area is the shoelace sum (scipy ConvexHull(...).volume of a convex
//   polygon), iou_2d = inter / (a1 + a2 - inter); iou_3d multiplies the BEV intersection by the overlap of the y extents
//   and divides by vol1 + vol2 - inter_vol with box3d_vol = |c0-c1| * |c1-c2| * |c0-c4|.
// Callers: get_3d_box / compute_box3d_iou (roi_seg_box3d_dataset.py:84-139: the iou2ds / iou3ds end points of
// semisup_v1_sunrgbd.get_iou_summary) and BoxPCFitDataset.perturb_box_to_diff_ious (box_pc_fit_dataset.py:211-244), the
// rejection sampler that draws perturbed boxes until their IoU with the original falls in a band -- a serial python
// loop per box there, one thread per box here with a counter-based Philox stream per (box, attempt).
// One thread per box pair; everything lives in registers / local arrays (a quad clipped by 4 half-planes has <= 8 vertices).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "simt_ops.cuh"     // philox4x32_10

namespace t3d {

// corners of get_3d_box (roi_seg_box3d_dataset.py:84-100): size = (l, w, h), R = roty(heading), upright camera coordinates
__device__ __forceinline__ void get_3d_box_dev(const float l, const float w, const float h, const float heading, const float cx,
                                               const float cy, const float cz, float (&c)[8][3]) {
  float s, co;
  sincosf(heading, &s, &co);
  const float xs[8] = {l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2, -l / 2};
  const float ys[8] = {h / 2, h / 2, h / 2, h / 2, -h / 2, -h / 2, -h / 2, -h / 2};
  const float zs[8] = {w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2};
#pragma unroll
  for (int i = 0; i < 8; ++i) {      // roty: [[c, 0, s], [0, 1, 0], [-s, 0, c]]
    c[i][0] = co * xs[i] + s * zs[i] + cx;
    c[i][1] = ys[i] + cy;
    c[i][2] = -s * xs[i] + co * zs[i] + cz;
  }
}

__device__ __forceinline__ float poly_area_dev(const float (*p)[2], int n) {   // 0.5 * |sum x_i y_{i-1} - y_i x_{i-1}|
  float a = 0.f;
  for (int i = 0; i < n; ++i) {
    const int j = (i + n - 1) % n;
    a += p[i][0] * p[j][1] - p[i][1] * p[j][0];
  }
  return 0.5f * fabsf(a);
}

// Sutherland-Hodgman (box_util.polygon_clip): subject clipped by every edge of the (convex) clip polygon; returns the
// vertex count (0: empty)
__device__ __forceinline__ int polygon_clip_dev(const float (*subject)[2], int ns, const float (*clip)[2], int nc, float (*out)[2]) {
  float a[16][2], b[16][2];
  int na = ns;
  for (int i = 0; i < ns; ++i) { a[i][0] = subject[i][0]; a[i][1] = subject[i][1]; }
  float cp1x = clip[nc - 1][0], cp1y = clip[nc - 1][1];
  for (int k = 0; k < nc; ++k) {
    const float cp2x = clip[k][0], cp2y = clip[k][1];
    int nb = 0;
    float sx = a[na - 1][0], sy = a[na - 1][1];
    auto inside = [&](float px, float py) { return (cp2x - cp1x) * (py - cp1y) > (cp2y - cp1y) * (px - cp1x); };
    for (int i = 0; i < na; ++i) {
      const float ex = a[i][0], ey = a[i][1];
      const bool ie = inside(ex, ey), is = inside(sx, sy);
      if (ie != is) {      // computeIntersection
        const float dcx = cp1x - cp2x, dcy = cp1y - cp2y, dpx = sx - ex, dpy = sy - ey;
        const float n1 = cp1x * cp2y - cp1y * cp2x, n2 = sx * ey - sy * ex, t1 = dcx * dpy, t2 = dcy * dpx, den = t1 - t2;
        // s-e (anti)parallel to the clip edge: both end points lie on the clip line up to round-off (coincident edges of
        // equal boxes) and the published formula divides by ~0; any point of the segment is the crossing -> take e
        const bool par = fabsf(den) <= 1e-6f * (fabsf(t1) + fabsf(t2));
        const float n3 = 1.0f / den;
        if (nb < 16) { b[nb][0] = par ? ex : (n1 * dpx - n2 * dcx) * n3; b[nb][1] = par ? ey : (n1 * dpy - n2 * dcy) * n3; ++nb; }
      }
      if (ie && nb < 16) { b[nb][0] = ex; b[nb][1] = ey; ++nb; }
      sx = ex; sy = ey;
    }
    cp1x = cp2x; cp1y = cp2y;
    if (nb == 0) return 0;
    na = nb;
    for (int i = 0; i < nb; ++i) { a[i][0] = b[i][0]; a[i][1] = b[i][1]; }
  }
  for (int i = 0; i < na; ++i) { out[i][0] = a[i][0]; out[i][1] = a[i][1]; }
  return na;
}

// box_util.box3d_iou on two corner sets
__device__ __forceinline__ void box3d_iou_dev(const float (&c1)[8][3], const float (&c2)[8][3], float& iou3d, float& iou2d) {
  float r1[4][2], r2[4][2], inter[16][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r1[i][0] = c1[3 - i][0]; r1[i][1] = c1[3 - i][2];
    r2[i][0] = c2[3 - i][0]; r2[i][1] = c2[3 - i][2];
  }
  const float area1 = poly_area_dev(r1, 4), area2 = poly_area_dev(r2, 4);
  const int n = polygon_clip_dev(r1, 4, r2, 4, inter);
  // the intersection cannot exceed either rectangle (guards the ratio against round-off in degenerate configurations)
  const float inter_area = n >= 3 ? fminf(poly_area_dev(inter, n), fminf(area1, area2)) : 0.f;
  iou2d = inter_area / (area1 + area2 - inter_area);
  const float ymax = fminf(c1[0][1], c2[0][1]), ymin = fmaxf(c1[4][1], c2[4][1]);
  const float inter_vol = inter_area * fmaxf(0.f, ymax - ymin);
  auto dist = [](const float* p, const float* q) {
    const float dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
    return sqrtf(dx * dx + dy * dy + dz * dz);
  };
  const float vol1 = dist(c1[0], c1[1]) * dist(c1[1], c1[2]) * dist(c1[0], c1[4]);
  const float vol2 = dist(c2[0], c2[1]) * dist(c2[1], c2[2]) * dist(c2[0], c2[4]);
  iou3d = inter_vol / (vol1 + vol2 - inter_vol);
}

__global__ void get_3d_box_kernel(const float* __restrict__ size, const float* __restrict__ heading, const float* __restrict__ center,
                                  int B, float* __restrict__ corners) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float c[8][3];
  get_3d_box_dev(size[b * 3], size[b * 3 + 1], size[b * 3 + 2], heading[b], center[b * 3], center[b * 3 + 1], center[b * 3 + 2], c);
  for (int i = 0; i < 8; ++i)
    for (int k = 0; k < 3; ++k) corners[(size_t)b * 24 + i * 3 + k] = c[i][k];
}

__global__ void box3d_iou_kernel(const float* __restrict__ corners1, const float* __restrict__ corners2, int B, float* __restrict__ iou3d,
                                 float* __restrict__ iou2d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float c1[8][3], c2[8][3];
  for (int i = 0; i < 8; ++i)
    for (int k = 0; k < 3; ++k) { c1[i][k] = corners1[(size_t)b * 24 + i * 3 + k]; c2[i][k] = corners2[(size_t)b * 24 + i * 3 + k]; }
  float i3, i2;
  box3d_iou_dev(c1, c2, i3, i2);
  if (iou3d) iou3d[b] = i3;
  if (iou2d) iou2d[b] = i2;
}

// roi_seg_box3d_dataset.compute_box3d_iou (:102-139): argmax-select (first max), class2angle (to_label_format wraps above pi),
// class2size (mean + residual), get_3d_box on prediction and label, box3d_iou
struct ComputeIouArgs {
  const float *center_pred, *heading_logits, *heading_residuals, *size_logits, *size_residuals;
  const float* center_label; const int* heading_class_label; const float* heading_residual_label;
  const int* size_class_label; const float* size_residual_label;
  const float* mean_size;       // [NS,3]
  int B, NH, NS;
  float *iou2ds, *iou3ds;
};
__device__ __forceinline__ float class2angle_dev(int cls, float residual, int num_class) {
  const float angle = (float)cls * (6.28318530717958647692f / (float)num_class) + residual;
  return angle > 3.14159265358979323846f ? angle - 6.28318530717958647692f : angle;
}
__global__ void compute_box3d_iou_kernel(const ComputeIouArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  int hc = 0, sc = 0;
  for (int j = 1; j < a.NH; ++j) if (a.heading_logits[b * a.NH + j] > a.heading_logits[b * a.NH + hc]) hc = j;
  for (int j = 1; j < a.NS; ++j) if (a.size_logits[b * a.NS + j] > a.size_logits[b * a.NS + sc]) sc = j;
  const float ha = class2angle_dev(hc, a.heading_residuals[b * a.NH + hc], a.NH);
  const float* sr = a.size_residuals + ((size_t)b * a.NS + sc) * 3;
  float c1[8][3], c2[8][3];
  get_3d_box_dev(a.mean_size[sc * 3] + sr[0], a.mean_size[sc * 3 + 1] + sr[1], a.mean_size[sc * 3 + 2] + sr[2], ha,
                 a.center_pred[b * 3], a.center_pred[b * 3 + 1], a.center_pred[b * 3 + 2], c1);
  const int hl = a.heading_class_label[b], sl = a.size_class_label[b];
  const float hal = class2angle_dev(hl, a.heading_residual_label[b], a.NH);
  get_3d_box_dev(a.mean_size[sl * 3] + a.size_residual_label[b * 3], a.mean_size[sl * 3 + 1] + a.size_residual_label[b * 3 + 1],
                 a.mean_size[sl * 3 + 2] + a.size_residual_label[b * 3 + 2], hal, a.center_label[b * 3], a.center_label[b * 3 + 1],
                 a.center_label[b * 3 + 2], c2);
  float i3, i2;
  box3d_iou_dev(c1, c2, i3, i2);
  a.iou2ds[b] = i2;
  a.iou3ds[b] = i3;
}

// BoxPCFitDataset.perturb_box_to_diff_ious (box_pc_fit_dataset.py:211-244) with a counter-based stream: attempt t of box b
// draws Philox4x32-10(counter = (t, 0|1, b, 0), key = seed): words 0..2 of block 0 -> centre deltas, word 3 -> angle delta,
// words 0..2 of block 1 -> size deltas; u = (word >> 8) * 2^-24 in [0, 1).  Accepts the first attempt whose 3D IoU with the
// original box lies strictly inside (lo, hi) (inrange :41-42); gives up after max_attempts (attempts = -1, last draw kept).
struct PerturbArgs {
  const float *center, *size, *heading;     // [B,3], [B,3] (l,w,h), [B]
  const float* bounds;                      // [B,2] IoU band per box
  int B, max_attempts;
  float center_perturbation, size_perturbation, angle_perturbation;
  unsigned long long seed;
  float *new_center, *new_size, *new_heading, *iou3d, *d_center, *d_size, *d_angle;
  int* attempts;
};
__device__ __forceinline__ float u01_dev(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }
__global__ void perturb_boxes_kernel(const PerturbArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const float lo = a.bounds[b * 2], hi = a.bounds[b * 2 + 1];
  const float scale = 1.0f - 0.5f * (lo + hi);                 // less perturbation if the band is close to 1
  const float cp = a.center_perturbation * scale, sp = a.size_perturbation * scale, ap = a.angle_perturbation * scale;
  const float cx = a.center[b * 3], cy = a.center[b * 3 + 1], cz = a.center[b * 3 + 2];
  const float l = a.size[b * 3], w = a.size[b * 3 + 1], h = a.size[b * 3 + 2], hd = a.heading[b];
  const uint32_t k0 = (uint32_t)(a.seed & 0xFFFFFFFFull), k1 = (uint32_t)(a.seed >> 32);
  float c0[8][3];
  get_3d_box_dev(l, w, h, hd, cx, cy, cz, c0);
  float dc[3] = {0, 0, 0}, ds[3] = {0, 0, 0}, da = 0.f, iou = -1.f;
  int t = 0;
  bool ok = false;
  for (; t < a.max_attempts && !ok; ++t) {
    uint32_t r0[4], r1[4];
    philox4x32_10((uint32_t)t, 0u, (uint32_t)b, 0u, k0, k1, r0);
    philox4x32_10((uint32_t)t, 1u, (uint32_t)b, 0u, k0, k1, r1);
#pragma unroll
    for (int k = 0; k < 3; ++k) dc[k] = -cp + (2.0f * cp) * u01_dev(r0[k]);
    da = ap * u01_dev(r0[3]);
    ds[0] = l * (-sp + (2.0f * sp) * u01_dev(r1[0]));
    ds[1] = w * (-sp + (2.0f * sp) * u01_dev(r1[1]));
    ds[2] = h * (-sp + (2.0f * sp) * u01_dev(r1[2]));
    float c1[8][3], i2;
    get_3d_box_dev(l + ds[0], w + ds[1], h + ds[2], hd + da, cx + dc[0], cy + dc[1], cz + dc[2], c1);
    box3d_iou_dev(c0, c1, iou, i2);
    ok = (iou > lo) && (iou < hi);
  }
  a.attempts[b] = ok ? t : -1;
  a.iou3d[b] = iou;
  a.new_heading[b] = hd + da;
  a.d_angle[b] = da;
  const float nc[3] = {cx + dc[0], cy + dc[1], cz + dc[2]}, nsz[3] = {l + ds[0], w + ds[1], h + ds[2]};
  for (int k = 0; k < 3; ++k) {
    a.new_center[b * 3 + k] = nc[k]; a.new_size[b * 3 + k] = nsz[k];
    a.d_center[b * 3 + k] = dc[k]; a.d_size[b * 3 + k] = ds[k];
  }
}

}  // namespace t3d
