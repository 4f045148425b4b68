// Frustum batch assembly (the input side of the hot path): ROISegBoxDataset.__getitem__ + get_batch
// (sunrgbd_detection/roi_seg_box3d_dataset.py:259-345, 370-417) for a whole batch in one launch.  The reference builds a
// batch with a python loop over frustums: np.random.choice resampling to `npoints`, the centre-view rotation
// (rotate_pc_along_y by pi/2 + frustum_angle, :346-368), the optional flip / shift augmentation and the label encoding
// (box centre, angle2class, size2class).  Here the dataset lives on the device as one flat point array + offsets; one CTA
// per batch slot gathers, rotates and augments the points, and its first thread encodes the labels in double precision
// (the reference's numpy arithmetic is float64).  The random draws (choice, flip, shifts) are inputs, so the caller can
// replay numpy's stream in the reference's order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace t3d {

struct AssembleArgs {
  const float* points; int C_src;     // [P, C_src] all frustums back to back
  const int* labels;                  // [P] or null
  const long long* pt_off;            // [F+1]
  const int* sel;                     // [B] frustum of each batch slot
  const int* choice;                  // [B,N] point index inside the frustum
  const float* frustum_angle;         // [F]
  const float* box3d;                 // [F,8,3] or null (rgb-detection layout has no labels)
  const float* heading;               // [F]
  const float* size;                  // [F,3] (l,w,h)
  const int* cls;                     // [F] class id (type2class)
  const float* mean_size;             // [NC,3]
  const unsigned char* flip;          // [B] or null
  const float* shift_z; const float* shift_y;   // [B] or null
  int B, N, C_out, rotate_to_center, NH;
  float* batch_data;                  // [B,N,C_out]
  int* batch_label;                   // [B,N] or null
  float* center;                      // [B,3]
  int* heading_class; float* heading_residual;    // [B]
  int* size_class; float* size_residual;          // [B], [B,3]
  float* rot_angle;                   // [B]
};

__global__ void __launch_bounds__(256) assemble_batch_kernel(const AssembleArgs a) {
  const int b = blockIdx.x;
  const int f = a.sel[b];
  const long long p0 = a.pt_off[f];
  const double rot_d = 1.5707963267948966 + (double)a.frustum_angle[f];        // get_center_view_rot_angle
  const float rot = (float)rot_d;
  float sn = 0.f, cs = 1.f;
  if (a.rotate_to_center) sincosf(rot, &sn, &cs);
  const bool flip = a.flip && a.flip[b];
  const float sz = a.shift_z ? a.shift_z[b] : 0.f, sy = a.shift_y ? a.shift_y[b] : 0.f;
  for (int n = threadIdx.x; n < a.N; n += 256) {
    const long long src = p0 + a.choice[(size_t)b * a.N + n];
    const float* p = a.points + src * a.C_src;
    float x = p[0], y = p[1], z = p[2];
    if (a.rotate_to_center) { const float xr = cs * x - sn * z, zr = sn * x + cs * z; x = xr; z = zr; }   // rotate_pc_along_y
    if (flip) x = -x;
    z += sz; y += sy;
    float* o = a.batch_data + ((size_t)b * a.N + n) * a.C_out;
    o[0] = x; o[1] = y; o[2] = z;
    for (int c = 3; c < a.C_out; ++c) o[c] = p[c];
    if (a.batch_label && a.labels) a.batch_label[(size_t)b * a.N + n] = a.labels[src];
  }
  if (threadIdx.x == 0) {
    a.rot_angle[b] = rot;
    if (a.box3d) {
      const float* bx = a.box3d + (size_t)f * 24;
      double cx = 0.5 * ((double)bx[0] + (double)bx[18]), cy = 0.5 * ((double)bx[1] + (double)bx[19]), cz = 0.5 * ((double)bx[2] + (double)bx[20]);
      double h = (double)a.heading[f];
      if (a.rotate_to_center) {
        const double c = cos(rot_d), s = sin(rot_d);
        const double xr = c * cx - s * cz, zr = s * cx + c * cz;
        cx = xr; cz = zr;
        h -= rot_d;
      }
      if (flip) { cx = -cx; h = 3.141592653589793 - h; }
      cz += (double)sz; cy += (double)sy;
      a.center[b * 3] = (float)cx; a.center[b * 3 + 1] = (float)cy; a.center[b * 3 + 2] = (float)cz;
      // angle2class (:47-62): python's % keeps the sign of the divisor
      const double two_pi = 6.283185307179586, per = two_pi / (double)a.NH;
      double ang = fmod(h, two_pi); if (ang < 0) ang += two_pi;
      double sh = fmod(ang + per / 2, two_pi); if (sh < 0) sh += two_pi;
      const int cid = (int)(sh / per);
      a.heading_class[b] = cid;
      a.heading_residual[b] = (float)(sh - (cid * per + per / 2));
      const int k = a.cls[f];
      a.size_class[b] = k;                                                     // size2class (:73-77)
      for (int j = 0; j < 3; ++j) a.size_residual[b * 3 + j] = (float)((double)a.size[f * 3 + j] - (double)a.mean_size[k * 3 + j]);
    }
  }
}

// Wire format -> placeholder layout: the frustum points cross PCIe as xyz fp32 [n,3] + rgb uint8 [n,3] (SUN-RGBD colours are
// im2double of 8-bit images, sunrgbd_data: rgb = k / 255) = 15 B / point instead of the 24 B of the fp32 (B,N,6) placeholder
// of semisup_v1_sunrgbd.py:39.  One thread per output float4 (stores fully coalesced, the reads hit L1): out[p][0:3] =
// xyz[p], out[p][3:6] = (float)rgb[p] / 255 (IEEE division: bit-identical to numpy's float32 k / 255).
__global__ void __launch_bounds__(256) assemble_points_kernel(const float* __restrict__ xyz, const uint8_t* __restrict__ rgb,
                                                              long long n_points, float* __restrict__ out) {
  const long long nq = (n_points * 6 + 3) / 4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long e = q * 4 + j;
      const long long p = e / 6;
      const int c = (int)(e - p * 6);
      v[j] = 0.0f;
      if (p < n_points) v[j] = c < 3 ? __ldg(xyz + p * 3 + c) : __fdiv_rn((float)__ldg(rgb + p * 3 + (c - 3)), 255.0f);
    }
    if (q * 4 + 3 < n_points * 6) {
      *reinterpret_cast<float4*>(out + q * 4) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      for (int j = 0; j < 4; ++j)
        if (q * 4 + j < n_points * 6) out[q * 4 + j] = v[j];
    }
  }
}

}  // namespace t3d
