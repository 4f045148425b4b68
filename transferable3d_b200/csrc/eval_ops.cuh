// Detection evaluation: the matching loop of eval_det.eval_det_cls (sunrgbd_detection/eval_det.py:
// 118-145) for one class.  The reference walks the detections in descending score order and, for each, computes the 3D
// IoU against every ground-truth box of the same image in python; a detection is a true positive if its best overlap
// exceeds the threshold and that ground-truth box has not been claimed yet.  The only sequential dependency is between
// detections of the SAME image, so one thread owns one image and walks that image's detections in score order; the
// oriented-box IoU is box_ops.cuh's box3d_iou_dev (the `get_iou` hook of eval_det.py:63-69).
#pragma once
#include "box_ops.cuh"

namespace t3d {

struct DetMatchArgs {
  const float* det_corners;   // [nd,8,3], detections of one class sorted by descending score
  const int* img_det_off;     // [nimg+1]  detections of image i = img_det_idx[off[i] .. off[i+1]) (ascending = score order)
  const int* img_det_idx;     // [nd]
  const float* gt_corners;    // [ng,8,3], grouped by image
  const int* img_gt_off;      // [nimg+1]
  int nimg;
  float ovthresh;
  float* tp; float* fp;       // [nd] 0 / 1
  float* ovmax; int* jmax;    // [nd] best overlap and the index (within the image) of its box, -inf / -1 without boxes
  unsigned char* gt_det;      // [ng] scratch: claimed flags (zero-initialised by the caller)
};

__global__ void det_match_kernel(const DetMatchArgs a) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= a.nimg) return;
  const int g0 = a.img_gt_off[img], g1 = a.img_gt_off[img + 1];
  for (int q = a.img_det_off[img]; q < a.img_det_off[img + 1]; ++q) {
    const int d = a.img_det_idx[q];
    float bb[8][3];
    for (int i = 0; i < 24; ++i) bb[i / 3][i % 3] = a.det_corners[(size_t)d * 24 + i];
    float best = -INFINITY; int jbest = -1;
    for (int j = g0; j < g1; ++j) {
      float gt[8][3];
      for (int i = 0; i < 24; ++i) gt[i / 3][i % 3] = a.gt_corners[(size_t)j * 24 + i];
      float i3, i2;
      box3d_iou_dev(bb, gt, i3, i2);
      if (i3 > best) { best = i3; jbest = j - g0; }        // strict: the first maximum wins (eval_det.py:131-133)
    }
    float tp = 0.f, fp = 0.f;
    if (best > a.ovthresh) {
      if (!a.gt_det[g0 + jbest]) { tp = 1.f; a.gt_det[g0 + jbest] = 1; }
      else fp = 1.f;
    } else {
      fp = 1.f;
    }
    a.tp[d] = tp; a.fp[d] = fp;
    if (a.ovmax) a.ovmax[d] = best;
    if (a.jmax) a.jmax[d] = jbest;
  }
}

}  // namespace t3d
