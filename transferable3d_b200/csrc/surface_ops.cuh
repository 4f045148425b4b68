// weak_losses.get_surface_loss (models/weak_losses.py:240-265) with the distance family it calls
// (models/tf_util.py: tf_create_3D_box_by_surface_centers :893-943, tf_distance_to_box_surfaces :610-677,
// tf_distance_to_closest_3D_box_surface(_multi) :679-720), forward and backward in one pass over the points.
//
// Per point p of frustum b, box (c, dims * scale, theta):  ray = p - c, r = |ray|, u = R^T ray (box frame:
// u_x = cos dx - sin dz, u_y = dy, u_z = sin dx + cos dz).  The six surfaces (x+, x-, y+, y-, z+, z-) have inward normals
// R(-+e_k), so  perp_i = ray . n_i = -+u_k  and  (p0 - l0) . n_i = -H_k  (H = l/2, h/2, w/2):
//     dist_center_to_surface_i = r * (-H_k) / (perp_i + 1e-5)        (H itself when the point sits on the centre)
//     dist_point_to_surface_i  = | r - dist_center_to_surface_i |
//     d = min_i dist_point_to_surface_i      -- the reference takes the min of the UNCLEANED distances (:707): surfaces
//                                              hit behind the centre and outside the box take part, replicated here
//     loss_b = mean_n max(0, d - margin) * soft_mask[b,n]            (the product uses soft_mask itself, so the mask
//                                              gradient flows whatever train_seg says (:247,:257) -- replicated)
// Gradients (hand-derived, one pass): f = r + r H / (q + eps) with q = perp_i, d = |f|;
//   df/dr = 1 + H/(q+eps), df/dH = r/(q+eps), df/dq = -r H/(q+eps)^2;  dr/dc = -ray/r;  q = sigma u_k, du/dc = -R^T,
//   du_x/dtheta = -u_z, du_z/dtheta = u_x;  dH/ddim = scale/2.  train_box = (center, dims, orient) flags gate the groups.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace t3d {

struct SurfaceLossArgs {
  const float* pc; int C;             // [B,N,C], xyz = first three channels
  const float* soft_mask;             // [B,N]
  const float* center; const float* dims; const float* orient;    // [B,3] (l,w,h) [B]
  int B, N;
  float margin, scale_dims;
  int train_center, train_dims, train_orient;
  const float* upstream;              // [B] d total / d loss_b, or null (forward only)
  float* loss;                        // [B]
  float* g_box;                       // [B,7] d total / d (center, dims, orient), or null
  float* g_mask;                      // [B,N] d total / d soft_mask, or null
};

__global__ void __launch_bounds__(256) surface_loss_kernel(const SurfaceLossArgs a) {
  const int b = blockIdx.x;
  const float cx = a.center[b * 3], cy = a.center[b * 3 + 1], cz = a.center[b * 3 + 2];
  const float H[3] = {0.5f * a.dims[b * 3] * a.scale_dims, 0.5f * a.dims[b * 3 + 2] * a.scale_dims,
                      0.5f * a.dims[b * 3 + 1] * a.scale_dims};      // half extents along box x (l), y (h), z (w)
  float sn, cs;
  sincosf(a.orient[b], &sn, &cs);
  const float up = a.upstream ? a.upstream[b] : 0.0f;
  const float invN = 1.0f / (float)a.N;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};         // loss, g cx cy cz, g l w h, g theta
  for (int n = threadIdx.x; n < a.N; n += 256) {
    const float* p = a.pc + ((size_t)b * a.N + n) * a.C;
    const float dx = p[0] - cx, dy = p[1] - cy, dz = p[2] - cz;
    const float r = sqrtf(dx * dx + dy * dy + dz * dz);
    const float u[3] = {cs * dx - sn * dz, dy, sn * dx + cs * dz};
    const bool at_centre = (fabsf(dx) + fabsf(dy) + fabsf(dz)) == 0.0f;
    float best = 3.0e38f, bf = 0.f, bq = 1.f; int bk = 0; float bs = 1.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int k = i >> 1;
      const float sg = (i & 1) ? 1.0f : -1.0f;                       // perp = sigma * u_k: x+ -> -u_x, x- -> +u_x, ...
      const float q = sg * u[k] + 1e-5f;
      const float dcs = at_centre ? H[k] : r * (-H[k]) / q;          // dist_center_to_surface
      const float f = r - dcs;
      const float d = fabsf(f);
      if (d < best) { best = d; bf = f; bq = q; bk = k; bs = sg; }   // first minimum (tf.reduce_min value; ties share it)
    }
    const float m = a.soft_mask[(size_t)b * a.N + n];
    const float over = best - a.margin;
    const float li = over > 0.0f ? over : 0.0f;
    acc[0] += li * m;
    if (a.upstream) {
      if (a.g_mask) a.g_mask[(size_t)b * a.N + n] = up * li * invN;
      if (over > 0.0f && a.g_box) {
        const float gf = up * m * invN * (bf > 0.0f ? 1.0f : (bf < 0.0f ? -1.0f : 0.0f));
        if (at_centre) {
          if (a.train_dims) acc[4 + (bk == 0 ? 0 : (bk == 1 ? 2 : 1))] += -gf * 0.5f * a.scale_dims;      // f = 0 - H
        } else {
          const float Hk = H[bk];
          const float df_dr = 1.0f + Hk / bq, df_dH = r / bq, df_dq = -r * Hk / (bq * bq);
          if (a.train_center) {
            // dr/dc = -ray/r ; dq/dc = sigma * d u_k / dc = -sigma * (row k of R^T)
            const float rk[3][3] = {{cs, 0.f, -sn}, {0.f, 1.f, 0.f}, {sn, 0.f, cs}};
            acc[1] += gf * (df_dr * (-dx / r) + df_dq * bs * (-rk[bk][0]));
            acc[2] += gf * (df_dr * (-dy / r) + df_dq * bs * (-rk[bk][1]));
            acc[3] += gf * (df_dr * (-dz / r) + df_dq * bs * (-rk[bk][2]));
          }
          if (a.train_dims) acc[4 + (bk == 0 ? 0 : (bk == 1 ? 2 : 1))] += gf * df_dH * 0.5f * a.scale_dims;
          if (a.train_orient) {
            const float duk = bk == 0 ? -u[2] : (bk == 2 ? u[0] : 0.0f);
            acc[7] += gf * df_dq * bs * duk;
          }
        }
      }
    }
  }
  __shared__ float sh[8][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
    if (threadIdx.x == 0) a.loss[b] = t * invN;
    else if (a.g_box && a.upstream) a.g_box[(size_t)b * 7 + threadIdx.x - 1] = t;
  }
}

// weak_losses.get_inactive_volume_loss_v1 (models/weak_losses.py:38-67): per trained class the mean over its samples of
// max(0, margin_c - l w h) (0 for an empty group), then the mean over the trained classes.  One CTA; class id = argmax of
// the one-hot row.  The value is written to out[0]; with w != 0 it is also accumulated the way get_semi_loss_final combines
// it (semisup_v1_sunrgbd.py:348-360, :393-398): total[4] (weak_loss) += w * iv, total[0] += mult * w * iv, and
// g_reg[b, 3:6] += mult * w * d iv / d dims_b.
struct InactiveVolArgs {
  const float* dims;        // [B,3] regression-format dims (l, w, h)
  const float* one_hot;     // [B,NC]
  const float* margins;     // [NC]
  int B, NC;
  unsigned train_mask;      // bit c: class c is in inactive_vol_train_classes
  float w, mult;
  float* out;               // [1]
  float* total;             // [8] of t3d_semi_loss, or null
  float* g_reg;             // [B,7] of t3d_semi_loss, or null
};
__global__ void __launch_bounds__(256) inactive_volume_kernel(const InactiveVolArgs a) {
  __shared__ float viol[32], cnt[32];
  if (threadIdx.x < 32) { viol[threadIdx.x] = 0.f; cnt[threadIdx.x] = 0.f; }
  __syncthreads();
  auto cls_of = [&](int b) {
    int cid = 0; float best = a.one_hot[(size_t)b * a.NC];
    for (int c = 1; c < a.NC; ++c) { const float v = a.one_hot[(size_t)b * a.NC + c]; if (v > best) { best = v; cid = c; } }
    return cid;
  };
  for (int b = threadIdx.x; b < a.B; b += 256) {
    const int c = cls_of(b);
    if (!((a.train_mask >> c) & 1u)) continue;
    const float vol = a.dims[b * 3] * a.dims[b * 3 + 1] * a.dims[b * 3 + 2];
    atomicAdd(&cnt[c], 1.0f);
    atomicAdd(&viol[c], fmaxf(0.0f, a.margins[c] - vol));
  }
  __syncthreads();
  const int ntrain = __popc(a.train_mask & ((a.NC >= 32) ? 0xffffffffu : ((1u << a.NC) - 1u)));
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int c = 0; c < a.NC; ++c) if (((a.train_mask >> c) & 1u) && cnt[c] > 0.f) s += viol[c] / cnt[c];
    const float iv = ntrain > 0 ? s / (float)ntrain : 0.f;
    a.out[0] = iv;
    if (a.total) { a.total[4] += a.w * iv; a.total[0] += a.mult * a.w * iv; }
  }
  __syncthreads();
  if (a.g_reg && a.w != 0.f && ntrain > 0) {
    for (int b = threadIdx.x; b < a.B; b += 256) {
      const int c = cls_of(b);
      if (!((a.train_mask >> c) & 1u)) continue;
      const float l = a.dims[b * 3], w = a.dims[b * 3 + 1], h = a.dims[b * 3 + 2];
      if (a.margins[c] - l * w * h > 0.0f) {
        const float g = -a.mult * a.w / (cnt[c] * (float)ntrain);
        a.g_reg[(size_t)b * 7 + 3] += g * w * h;
        a.g_reg[(size_t)b * 7 + 4] += g * l * h;
        a.g_reg[(size_t)b * 7 + 5] += g * l * w;
      }
    }
  }
}

}  // namespace t3d
