// extern "C" entry points of libt3d_b200.so (declared in include/t3d_b200.h).
#include <cmath>
#include <cstdlib>
#include "../../include/t3d_b200.h"
#include "common.cuh"
#include "simt_ops.cuh"
#include "chain_max.cuh"
#include "seg_stage2_pair.cuh"
#include "chain_x2.cuh"
#include "seg_stage2_x2.cuh"
#include "train_ops.cuh"
#include "loss_ops.cuh"
#include "box_ops.cuh"
#include "post_ops.cuh"
#include "eval_ops.cuh"
#include "surface_ops.cuh"
#include "data_ops.cuh"
#include "skinny_gemm.cuh"
#define T3D_SGEMM_WITH_EPILOGUES
#include "sgemm.cuh"
#define T3D_XGEMM_WITH_EPILOGUES
#include "xgemm.cuh"

using namespace t3d;

#define T3D_CHECK_LAUNCH()                          \
  do {                                              \
    cudaError_t e__ = cudaGetLastError();           \
    if (e__ != cudaSuccess) return (int)e__;        \
  } while (0)
#define T3D_CUDA(x)                                 \
  do {                                              \
    cudaError_t e__ = (x);                          \
    if (e__ != cudaSuccess) return (int)e__;        \
  } while (0)

static unsigned long long* g_trace = nullptr;
extern "C" int t3d_set_trace_buffer(void* dev_buf) { g_trace = reinterpret_cast<unsigned long long*>(dev_buf); return 0; }

static inline cudaStream_t S(t3d_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" int t3d_version(void) { return T3D_VERSION; }

extern "C" const char* t3d_error_string(int code) {
  if (code == 0) return "ok";
  if (code == T3D_ERR_ARG) return "t3d: invalid argument";
  if (code == T3D_ERR_SHAPE) return "t3d: unsupported shape";
  if (code == T3D_ERR_ALIGN) return "t3d: misaligned pointer";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "t3d: unknown error";
}

// fp32 GEMM engine for tile-sized problems: 1 = tcgen05 "bf16 x 3" split (xgemm.cuh: six products per MAC, fp32-accurate,
// default), 0 = CUDA-core SGEMM (sgemm.cuh), 2 = one bf16 product per MAC (approximate), 3 = "bf16 x 2" split (three
// products per MAC, operand error <= 2^-18: 30 x tighter than TF32).  Process-wide configuration, not per-call state.
static int g_f32_engine = 1;
static inline int xg_parts() { return g_f32_engine == 1 ? 3 : (g_f32_engine == 3 ? 2 : 1); }
// launches `CALL` (an expression using the template parameter P) for the engine's number of operand pieces
#define XG_BY_PARTS(CALL)                            \
  do {                                               \
    const int parts_ = xg_parts();                   \
    if (parts_ == 3) { constexpr int P = 3; CALL; }  \
    else if (parts_ == 2) { constexpr int P = 2; CALL; } \
    else { constexpr int P = 1; CALL; }              \
  } while (0)
extern "C" int t3d_set_f32_engine(int engine) {
  if (engine < 0 || engine > 3) return T3D_ERR_ARG;
  g_f32_engine = engine;
  return 0;
}
extern "C" int t3d_get_f32_engine(void) { return g_f32_engine; }

template <typename Kern>
static int xg_prepare(Kern kern) {      // opt in to 97 KB of dynamic shared memory (2 CTAs / SM)
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kXgSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  return (int)e;
}
// (M in [64, 128) with a long reduction -- the weight gradient of a 64-channel layer over B*N rows -- runs as a half-empty row
// tile: 0.42 ms against 1.39 ms on the CUDA-core kernel for [64 x 512] over 524288 rows, engine tc2)
static inline bool xg_fits(int M, int N, int K) { return g_f32_engine != 0 && (M >= 128 || (M >= 64 && K >= 4096)) && N >= 64 && K >= 32; }

// B_PRE path: forward / dgrad with the small operand pre-split once per call into the caller's workspace (xgemm.cuh)
static inline bool xg_pre_ok(int M, int N, int K, const void* ws, size_t ws_bytes) {
  return ws != nullptr && (((uintptr_t)ws) & 15) == 0 && M >= 4096 && K <= kXgMaxKChunk && ws_bytes >= xg_pre_bytes(N, K);
}

// Persistent forward / dgrad kernel (xg_pp_kernel) for K <= 128: with 4 stages or fewer per tile the per-CTA overheads and
// the epilogue dominate the one-tile kernel (measured: 98 / 117 / 140 against 86 / 103 / 121 TFLOP/s on the 128 -> 128 /
// 256 / 1024 forward layers); for longer K the two co-resident one-tile CTAs win (176 against 158 on dgrad 256 <- 512).
// t3d_linear_f32 never uses it: with an output to store its 4 epilogue warps are the limit (fp32-mode pipeline 43.6 k ->
// 24.8 k frustums/s with it everywhere, 32.7 k on the max-only layers alone).
static bool xg_use_pp(int K) { return K <= 128; }
// A-stationary kernel (xg_as_kernel): K <= 128 and at least four column tiles; for t3d_linear_f32 only the max-only layers
// (no Y to store: its 4 epilogue warps are the limit when every tile writes 64 KB).
static bool xg_use_as(int K, int ntn, bool linear = false, bool has_y = true) { return K <= 128 && ntn >= 4 && !(linear && has_y); }
// Function attributes (dynamic shared memory opt-in) and the SM count are per DEVICE: caches are keyed by the current one.
constexpr int kMaxDevices = 64;
static int cur_device() { int dev = 0; cudaGetDevice(&dev); return (dev >= 0 && dev < kMaxDevices) ? dev : 0; }
static int xg_num_sms() {
  static int n[kMaxDevices] = {0};
  const int dev = cur_device();
  if (n[dev] == 0) { int v = 148; cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); n[dev] = v; }
  return n[dev];
}
// runs `prep` once per device; returns its error code (0 = prepared)
template <typename F>
static int once_per_device(int (&state)[kMaxDevices], F prep) {
  const int dev = cur_device();
  if (state[dev] == 0) { const int e = prep(); if (e != 0) return e; state[dev] = 1; }
  return 0;
}
template <int PARTS, typename Kern>
static int xg_prepare_pp(Kern kern) {
  return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XgPP<PARTS>::SMEM);
}

extern "C" size_t t3d_gemm_ws_bytes(int N, int K) { return (N > 0 && K > 0) ? xg_pre_bytes(N, K) : 0; }

extern "C" int t3d_linear_f32_ws(const float* X, int ldx, const float* W, int ldw, const float* bias, const float* gbias,
                                 int rows_per_group, float* Y, int ldy, int M, int K, int N, int act, const float* rowmask,
                                 float* gmax, void* ws, size_t ws_bytes, t3d_stream_t stream) {
  if (!X || !W || (!Y && !gmax)) return T3D_ERR_ARG;
  if (M <= 0 || K <= 0 || N <= 0 || act < 0 || act > 3) return T3D_ERR_SHAPE;
  if ((gbias || gmax) && rows_per_group <= 0) return T3D_ERR_SHAPE;
  LinearArgs a{X, ldx, W, ldw, bias, gbias, rows_per_group, Y, ldy, M, K, N, act, rowmask, gmax};
  if (M >= 4096 && Y && !gbias && !rowmask && !gmax) {       // HBM-bound first / last layers (skinny_gemm.cuh)
    auto al16 = [](const void* p) { return (((uintptr_t)p) & 15) == 0; };
    if (K <= kSkinnyMax && N % 4 == 0 && N <= 1024 && 256 % (N / 4) == 0 && al16(Y) && ldy % 4 == 0 &&
        sizeof(float) * (size_t)(K * N + N) <= kSkinnySmemMax) {
      skinny_k_kernel<<<xg_num_sms() * 8, 256, sizeof(float) * (size_t)(K * N + N), S(stream)>>>(X, ldx, W, ldw, bias, Y, ldy, M, N, K, act);
      T3D_CHECK_LAUNCH();
      return 0;
    }
    if (N <= kSkinnyMax && act == 0 && K % 4 == 0 && K <= 1024 && al16(X) && ldx % 4 == 0 &&
        sizeof(float) * (size_t)N * K <= kSkinnySmemMax) {
      skinny_n_kernel<<<xg_num_sms() * 8, 256, sizeof(float) * (size_t)N * K, S(stream)>>>(X, ldx, W, 1, bias, Y, ldy, M, N, K, ldw);
      T3D_CHECK_LAUNCH();
      return 0;
    }
  }
  if (xg_fits(M, N, K)) {                    // tensor cores, bf16 x 3 split (xgemm.cuh)
    static int prepared[kMaxDevices] = {0};
    if (int e = once_per_device(prepared, [] {
          return xg_prepare(xlinear_kernel<3>) | xg_prepare(xlinear_kernel<2>) | xg_prepare(xlinear_kernel<1>) |
                 xg_prepare(xlinear_pre_kernel<3>) | xg_prepare(xlinear_pre_kernel<2>) | xg_prepare(xlinear_pre_kernel<1>) |
                 xg_prepare_pp<3>(xg_as_kernel<3, true>) | xg_prepare_pp<2>(xg_as_kernel<2, true>) | xg_prepare_pp<1>(xg_as_kernel<1, true>);
        }))
      return e;
    const int ntm = (M + kXgBM - 1) / kXgBM, ntn = (N + kXgBN - 1) / kXgBN;
    const int parts = xg_parts();
    XgOperands o{X, ldx, W, ldw, M, N, K, (K + kXgBK - 1) / kXgBK * kXgBK, xg_aligned16(X, ldx) ? 1 : 0, 0, ntn, g_trace, nullptr, 0};
    const dim3 grid((unsigned)ntm * ntn, 1);
    // CTA-pair kernel (256 x 256 tiles, no pre-split pass) for the engines with at most two operand pieces, when the grid
    // fills the machine and the A-stationary kernel does not apply (measured in the cfg5 step, where the frozen branches run
    // through this call: 15.5 -> 15.3 ms; restricted to K > 128: 15.4)
    const int ntm2 = (M + 255) / 256, ntn2 = (N + 255) / 256;
    if (parts <= 2 && M >= 256 && N >= 256 && (long long)ntm2 * ntn2 * 2 >= xg_num_sms() &&
        !(xg_pre_ok(M, N, K, ws, ws_bytes) && xg_use_as(K, ntn, true, Y != nullptr))) {
      static int prepared_pair[kMaxDevices] = {0};
      if (int e = once_per_device(prepared_pair, [] { return xg_prepare(xlinear_pair_kernel<2>) | xg_prepare(xlinear_pair_kernel<1>); })) return e;
      o.ntn = ntn2;
      const dim3 pgrid((unsigned)ntm2 * ntn2 * 2, 1);
      if (parts == 2) xlinear_pair_kernel<2><<<pgrid, kXgThreads, kXgSmemBytes, S(stream)>>>(a, o);
      else xlinear_pair_kernel<1><<<pgrid, kXgThreads, kXgSmemBytes, S(stream)>>>(a, o);
      T3D_CHECK_LAUNCH();
      return 0;
    }
    if (xg_pre_ok(M, N, K, ws, ws_bytes)) {
      o.bpre = reinterpret_cast<const uint8_t*>(ws);
      o.nkb = (K + 63) / 64;
      xg_presplit_kernel<<<dim3(o.nkb, ntn, 4), 256, 0, S(stream)>>>(W, ldw, 0, N, K, parts, reinterpret_cast<uint8_t*>(ws));
      if (xg_use_as(K, ntn, true, Y != nullptr)) {
        const int g = ntm < xg_num_sms() ? ntm : xg_num_sms();
        XG_BY_PARTS((xg_as_kernel<P, true><<<g, XgPP<P>::THREADS, XgPP<P>::SMEM, S(stream)>>>(GemmArgs{}, a, o)));
      } else XG_BY_PARTS((xlinear_pre_kernel<P><<<grid, kXgThreads, kXgSmemBytes, S(stream)>>>(a, o)));
    } else {
      XG_BY_PARTS((xlinear_kernel<P><<<grid, kXgThreads, kXgSmemBytes, S(stream)>>>(a, o)));
    }
    T3D_CHECK_LAUNCH();
    return 0;
  }
  if (M >= 128 && N >= 96 && K >= 16) {      // 128 x 128 tiles (sgemm.cuh)
    SgemmOperands o{X, ldx, 1, W, ldw, 1, M, N, K, (K + kSgBK - 1) / kSgBK * kSgBK, sg_aligned16(X, ldx) ? 1 : 0, sg_aligned16(W, ldw) ? 1 : 0};
    dim3 grid((M + kSgBM - 1) / kSgBM, (N + kSgBN - 1) / kSgBN);
    linear128_kernel<<<grid, 256, 0, S(stream)>>>(a, o);
    T3D_CHECK_LAUNCH();
    return 0;
  }
  dim3 grid((M + 63) / 64, (N + 63) / 64);
  linear_f32_kernel<<<grid, 256, 0, S(stream)>>>(a);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_linear_f32(const float* X, int ldx, const float* W, int ldw, const float* bias, const float* gbias,
                              int rows_per_group, float* Y, int ldy, int M, int K, int N, int act, const float* rowmask,
                              float* gmax, t3d_stream_t stream) {
  return t3d_linear_f32_ws(X, ldx, W, ldw, bias, gbias, rows_per_group, Y, ldy, M, K, N, act, rowmask, gmax, nullptr, 0, stream);
}

extern "C" int t3d_mask_centroid(const float* logits, const float* pc, int B, int N, int C, float* mask, int* count,
                                 float* mean, float* xyz_stage1, int* idx, t3d_stream_t stream) {
  if (!logits || !pc) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0 || C < 3) return T3D_ERR_SHAPE;
  if ((uintptr_t)logits & 7) return T3D_ERR_ALIGN;
  mask_centroid_kernel<<<B, 256, 0, S(stream)>>>(logits, pc, N, C, mask, count, mean, xyz_stage1, idx);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_resample(const int* idx, const int* count, int B, int N, int npoints, int mode, uint64_t seed,
                            const int* choice, int* indices, const float* pc, int C, const float* mean, int c_out,
                            float* object_pc, t3d_stream_t stream) {
  if (!idx || !count || !indices) return T3D_ERR_ARG;
  if (mode != 0 && mode != 1) return T3D_ERR_ARG;
  if (mode == 1 && !choice) return T3D_ERR_ARG;
  if (object_pc && (!pc || !mean || c_out < 3 || c_out > C)) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0 || N > 2048 || npoints <= 0 || npoints > 2048) return T3D_ERR_SHAPE;
  resample_kernel<<<B, kResampleThreads, 0, S(stream)>>>(idx, count, N, npoints, mode, (unsigned long long)seed, choice, indices, pc, C,
                                             mean, c_out, object_pc);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_build_tiles(const int* count, int B, int tile_pts, void* tiles, int* num_tiles, t3d_stream_t stream) {
  if (!count || !tiles || !num_tiles) return T3D_ERR_ARG;
  if (B <= 0 || tile_pts <= 0) return T3D_ERR_SHAPE;
  build_tiles_kernel<<<1, 1024, 0, S(stream)>>>(count, B, tile_pts, reinterpret_cast<int4*>(tiles), num_tiles);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_prepare_xyz(const float* pc, int B, int N, int C, const float* center, float* out, t3d_stream_t stream) {
  if (!pc || !out) return T3D_ERR_ARG;
  const size_t n = (size_t)B * N;
  prepare_xyz_kernel<<<(unsigned)((n + 255) / 256), 256, 0, S(stream)>>>(pc, B, N, C, center, out);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_boxpc_features(const float* pc, int B, int N, int C, const float* center, const float* dims,
                                  const float* orient, float* out, t3d_stream_t stream) {
  if (!pc || !center || !dims || !orient || !out) return T3D_ERR_ARG;
  const size_t n = (size_t)B * N;
  boxpc_features_kernel<<<(unsigned)((n + 255) / 256), 256, 0, S(stream)>>>(pc, B, N, C, center, dims, orient, out);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_parse_box(const t3d_parse_args* p, t3d_stream_t stream) {
  if (!p || !p->output || !p->mean_size || !p->orient_anchors) return T3D_ERR_ARG;
  if (p->reg_center && (!p->reg_dims || !p->reg_orient)) return T3D_ERR_ARG;
  if (p->B <= 0 || p->NH <= 0 || p->NS <= 0) return T3D_ERR_SHAPE;
  ParseArgs a{p->output, p->stage1_center, p->mean_size, p->orient_anchors, p->B, p->NH, p->NS,
              p->center, p->heading_scores, p->heading_res_norm, p->heading_res, p->size_scores, p->size_res_norm,
              p->size_res, p->reg_center, p->reg_dims, p->reg_orient};
  parse_box_kernel<<<(p->B + 127) / 128, 128, 0, S(stream)>>>(a);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_anchor_to_reg(const float* center, const float* dims_cls, const float* dims_reg, const float* orient_cls,
                                 const float* orient_reg, const float* dims_anchors, const float* orient_anchors, int B, int NS,
                                 int NH, float* out_center, float* out_dims, float* out_orient, t3d_stream_t stream) {
  if (!center || !dims_cls || !dims_reg || !orient_cls || !orient_reg || !dims_anchors || !orient_anchors || !out_center ||
      !out_dims || !out_orient)
    return T3D_ERR_ARG;
  anchor_to_reg_kernel<<<(B + 127) / 128, 128, 0, S(stream)>>>(center, dims_cls, dims_reg, orient_cls, orient_reg, dims_anchors,
                                                               orient_anchors, B, NS, NH, out_center, out_dims, out_orient);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_boxpc_refine(const t3d_refine_args* p, t3d_stream_t stream) {
  if (!p || !p->out9) return T3D_ERR_ARG;
  if (p->delta_center && !p->delta_size) return T3D_ERR_ARG;
  if (p->box_center && (!p->box_dims || !p->box_orient)) return T3D_ERR_ARG;
  if (p->tot_center && (!p->tot_size || !p->tot_angle)) return T3D_ERR_ARG;
  RefineArgs a{p->out9, p->B, p->weigh_pred_by_conf, p->weigh_during_test, p->fit_logits, p->fit_prob, p->pred_fit,
               p->delta_center, p->delta_size, p->delta_angle, p->box_center, p->box_dims, p->box_orient,
               p->tot_center, p->tot_size, p->tot_angle};
  boxpc_refine_kernel<<<(p->B + 127) / 128, 128, 0, S(stream)>>>(a);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_f2(const float* f_center, const float* f_hres, const float* f_sres, const float* tot_center,
                      const float* tot_angle, const float* tot_size, int B, int NH, int NS, float* f2_center, float* f2_hres,
                      float* f2_sres, t3d_stream_t stream) {
  if (!f_center || !f_hres || !f_sres || !tot_center || !tot_angle || !tot_size || !f2_center || !f2_hres || !f2_sres)
    return T3D_ERR_ARG;
  f2_kernel<<<(B + 127) / 128, 128, 0, S(stream)>>>(f_center, f_hres, f_sres, tot_center, tot_angle, tot_size, B, NH, NS,
                                                    f2_center, f2_hres, f2_sres);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_box3d_corners_helper(const float* centers, const float* headings, const float* sizes, int n, float* out,
                                        t3d_stream_t stream) {
  if (!centers || !headings || !sizes || !out) return T3D_ERR_ARG;
  box3d_corners_helper_kernel<<<(n + 127) / 128, 128, 0, S(stream)>>>(centers, headings, sizes, n, out);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_box3d_corners_all(const float* center, const float* heading_res, const float* size_res,
                                     const float* mean_size, const float* orient_anchors, int B, int NH, int NS, float* out,
                                     t3d_stream_t stream) {
  if (!center || !heading_res || !size_res || !mean_size || !orient_anchors || !out) return T3D_ERR_ARG;
  const int n = B * NH * NS;
  box3d_corners_all_kernel<<<(n + 127) / 128, 128, 0, S(stream)>>>(center, heading_res, size_res, mean_size, orient_anchors, B,
                                                                   NH, NS, out);
  T3D_CHECK_LAUNCH();
  return 0;
}

// ----------------------------------------------------------------------------- oriented 3D IoU / BoxPC perturbation
extern "C" int t3d_get_3d_box(const float* size, const float* heading, const float* center, int B, float* corners, t3d_stream_t stream) {
  if (!size || !heading || !center || !corners) return T3D_ERR_ARG;
  if (B <= 0) return T3D_ERR_SHAPE;
  get_3d_box_kernel<<<(B + 127) / 128, 128, 0, S(stream)>>>(size, heading, center, B, corners);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_box3d_iou(const float* corners1, const float* corners2, int B, float* iou3d, float* iou2d, t3d_stream_t stream) {
  if (!corners1 || !corners2 || (!iou3d && !iou2d)) return T3D_ERR_ARG;
  if (B <= 0) return T3D_ERR_SHAPE;
  box3d_iou_kernel<<<(B + 127) / 128, 128, 0, S(stream)>>>(corners1, corners2, B, iou3d, iou2d);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_compute_box3d_iou(const t3d_compute_iou_args* a, t3d_stream_t stream) {
  if (!a || !a->center_pred || !a->heading_logits || !a->heading_residuals || !a->size_logits || !a->size_residuals ||
      !a->center_label || !a->heading_class_label || !a->heading_residual_label || !a->size_class_label ||
      !a->size_residual_label || !a->mean_size || !a->iou2ds || !a->iou3ds)
    return T3D_ERR_ARG;
  if (a->B <= 0 || a->NH <= 0 || a->NS <= 0) return T3D_ERR_SHAPE;
  ComputeIouArgs k{a->center_pred, a->heading_logits, a->heading_residuals, a->size_logits, a->size_residuals, a->center_label,
                   a->heading_class_label, a->heading_residual_label, a->size_class_label, a->size_residual_label, a->mean_size,
                   a->B, a->NH, a->NS, a->iou2ds, a->iou3ds};
  compute_box3d_iou_kernel<<<(a->B + 127) / 128, 128, 0, S(stream)>>>(k);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_perturb_boxes(const t3d_perturb_args* a, t3d_stream_t stream) {
  if (!a || !a->center || !a->size || !a->heading || !a->bounds || !a->new_center || !a->new_size || !a->new_heading ||
      !a->iou3d || !a->d_center || !a->d_size || !a->d_angle || !a->attempts)
    return T3D_ERR_ARG;
  if (a->B <= 0 || a->max_attempts <= 0) return T3D_ERR_SHAPE;
  PerturbArgs k{a->center, a->size, a->heading, a->bounds, a->B, a->max_attempts, a->center_perturbation, a->size_perturbation,
                a->angle_perturbation, (unsigned long long)a->seed, a->new_center, a->new_size, a->new_heading, a->iou3d,
                a->d_center, a->d_size, a->d_angle, a->attempts};
  perturb_boxes_kernel<<<(a->B + 63) / 64, 64, 0, S(stream)>>>(k);
  T3D_CHECK_LAUNCH();
  return 0;
}

// ----------------------------------------------------------------------------- inference post-processing
extern "C" int t3d_inference_scores(const t3d_infer_score_args* a, t3d_stream_t stream) {
  if (!a || !a->logits || !a->heading_scores || !a->heading_residuals || !a->size_scores || !a->size_residuals || !a->heading_cls ||
      !a->heading_res || !a->size_cls || !a->size_res || !a->scores)
    return T3D_ERR_ARG;
  if (a->B <= 0 || a->N <= 0 || a->NH <= 0 || a->NS <= 0) return T3D_ERR_SHAPE;
  if ((uintptr_t)a->logits & 7) return T3D_ERR_ALIGN;
  InferScoreArgs k{a->logits, a->heading_scores, a->heading_residuals, a->size_scores, a->size_residuals, a->fit_prob, a->B, a->N,
                   a->NH, a->NS, a->pred_seg, a->mask_mean_prob, a->heading_cls, a->heading_res, a->size_cls, a->size_res, a->scores};
  inference_scores_kernel<<<a->B, 256, 0, S(stream)>>>(k);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_prediction_to_label(const float* center, const int* heading_cls, const float* heading_res, const int* size_cls,
                                       const float* size_res, const float* rot_angle, const float* mean_size, int B, int NH,
                                       float* out7, t3d_stream_t stream) {
  if (!center || !heading_cls || !heading_res || !size_cls || !size_res || !rot_angle || !mean_size || !out7) return T3D_ERR_ARG;
  if (B <= 0 || NH <= 0) return T3D_ERR_SHAPE;
  prediction_to_label_kernel<<<(B + 127) / 128, 128, 0, S(stream)>>>(center, heading_cls, heading_res, size_cls, size_res, rot_angle,
                                                                     mean_size, B, NH, out7);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_surface_loss(const t3d_surface_loss_args* a, t3d_stream_t stream) {
  if (!a || !a->pc || !a->soft_mask || !a->center || !a->dims || !a->orient || !a->loss) return T3D_ERR_ARG;
  if (a->B <= 0 || a->N <= 0 || a->C < 3) return T3D_ERR_SHAPE;
  SurfaceLossArgs k{a->pc, a->C, a->soft_mask, a->center, a->dims, a->orient, a->B, a->N, a->margin, a->scale_dims, a->train_center,
                    a->train_dims, a->train_orient, a->upstream, a->loss, a->g_box, a->g_mask};
  surface_loss_kernel<<<a->B, 256, 0, S(stream)>>>(k);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_inactive_volume_loss(const float* dims, const float* one_hot, const float* margins, int B, int NC,
                                        unsigned train_mask, float w, float mult, float* out, float* total, float* g_reg,
                                        t3d_stream_t stream) {
  if (!dims || !one_hot || !margins || !out) return T3D_ERR_ARG;
  if (B <= 0 || NC <= 0 || NC > 32) return T3D_ERR_SHAPE;
  InactiveVolArgs k{dims, one_hot, margins, B, NC, train_mask, w, mult, out, total, g_reg};
  inactive_volume_kernel<<<1, 256, 0, S(stream)>>>(k);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_assemble_frustum_batch(const t3d_assemble_args* a, t3d_stream_t stream) {
  if (!a || !a->points || !a->pt_off || !a->sel || !a->choice || !a->frustum_angle || !a->batch_data || !a->rot_angle) return T3D_ERR_ARG;
  if (a->box3d && (!a->heading || !a->size || !a->cls || !a->mean_size || !a->center || !a->heading_class || !a->heading_residual ||
                   !a->size_class || !a->size_residual))
    return T3D_ERR_ARG;
  if (a->B <= 0 || a->N <= 0 || a->C_src < 3 || a->C_out < 3 || a->C_out > a->C_src || a->NH <= 0) return T3D_ERR_SHAPE;
  AssembleArgs k{a->points, a->C_src, a->labels, a->pt_off, a->sel, a->choice, a->frustum_angle, a->box3d, a->heading, a->size, a->cls,
                 a->mean_size, a->flip, a->shift_z, a->shift_y, a->B, a->N, a->C_out, a->rotate_to_center, a->NH, a->batch_data,
                 a->batch_label, a->center, a->heading_class, a->heading_residual, a->size_class, a->size_residual, a->rot_angle};
  assemble_batch_kernel<<<a->B, 256, 0, S(stream)>>>(k);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_assemble_points(const float* xyz, const uint8_t* rgb, long long n_points, float* out, t3d_stream_t stream) {
  if (!xyz || !rgb || !out) return T3D_ERR_ARG;
  if (n_points <= 0) return T3D_ERR_SHAPE;
  if ((uintptr_t)out & 15) return T3D_ERR_ALIGN;
  assemble_points_kernel<<<xg_num_sms() * 8, 256, 0, S(stream)>>>(xyz, rgb, n_points, out);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_det_match(const t3d_det_match_args* a, t3d_stream_t stream) {
  if (!a || !a->det_corners || !a->img_det_off || !a->img_det_idx || !a->img_gt_off || !a->tp || !a->fp || !a->gt_det) return T3D_ERR_ARG;
  if (a->nimg <= 0 || a->nd <= 0 || a->ng < 0 || (a->ng > 0 && !a->gt_corners)) return T3D_ERR_SHAPE;
  T3D_CUDA(cudaMemsetAsync(a->gt_det, 0, (size_t)(a->ng > 0 ? a->ng : 1), S(stream)));
  DetMatchArgs k{a->det_corners, a->img_det_off, a->img_det_idx, a->gt_corners, a->img_gt_off, a->nimg, a->ovthresh, a->tp, a->fp,
                 a->ovmax, a->jmax, a->gt_det};
  det_match_kernel<<<(a->nimg + 63) / 64, 64, 0, S(stream)>>>(k);
  T3D_CHECK_LAUNCH();
  return 0;
}

// ----------------------------------------------------------------------------- tcgen05 chains
constexpr int kPackMax = 64;
struct PackTable { PackDesc d[kPackMax]; };
__global__ void __launch_bounds__(256) pack_table_kernel(const PackTable tab, uint8_t* arena) {
  const PackDesc d = tab.d[blockIdx.x];
  uint8_t* dst = arena + (size_t)blockIdx.x * kChunkBytes;
  for (int e = threadIdx.x; e < 128 * 64; e += 256) {
    const int r = e >> 6, kk = e & 63;
    float v = 0.0f;
    if (r < d.nrows && d.k0 + kk < d.k_total) v = d.W[(size_t)(d.k0 + kk) * d.ldw + d.row0 + r];
    uint8_t* o = dst + sw128_offset(r, kk >> 3) + (kk & 7) * 2;
    if (d.part == 0) {
      *reinterpret_cast<__nv_bfloat16*>(o) = __float2bfloat16_rn(v);
    } else {                                   // fp16 hi / lo images of W * scale (round to nearest both)
      v *= d.scale;
      const __half hi = __float2half_rn(v);
      *reinterpret_cast<__half*>(o) = d.part == 1 ? hi : __float2half_rn(v - __half2float(hi));
    }
  }
}
__global__ void scale_copy_kernel(float* dst, const float* src, int n, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] * scale;
}
__global__ void write4_kernel(float* dst, float a, float b, float c, float d) { dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d; }
static int scale_copy(float* dst, const float* src, int n, float scale, cudaStream_t st) {
  scale_copy_kernel<<<(n + 255) / 256, 256, 0, st>>>(dst, src, n, scale);
  T3D_CHECK_LAUNCH();
  return 0;
}

template <int KIND>
static int pack_chain_impl(const float* const* W, const float* const* bias, uint8_t* arena, cudaStream_t st) {
  using Sp = ChainSpec<KIND>;
  PackTable tab;
  int n = 0;
  for (int l = 0; l < Sp::NH; ++l)
    for (int nb = 0; nb < (Sp::HN(l) + 127) / 128; ++nb)
      for (int kb = 0; kb < Sp::HK(l) / 64; ++kb)
        tab.d[n++] = PackDesc{W[1 + l], Sp::HN(l), Sp::HK(l), kb * 64, nb * 128, Sp::HN(l) - nb * 128 < 128 ? Sp::HN(l) - nb * 128 : 128};
  for (int mt = 0; mt < Sp::FC / 128; ++mt)
    for (int kb = 0; kb < Sp::FK / 64; ++kb)
      tab.d[n++] = PackDesc{W[1 + Sp::NH], Sp::FC, Sp::FK, kb * 64, mt * 128, 128};
  if (n != chain_num_chunks<Sp>() || n > kPackMax) return T3D_ERR_SHAPE;
  pack_table_kernel<<<n, 256, 0, st>>>(tab, arena);
  T3D_CHECK_LAUNCH();
  float* f = reinterpret_cast<float*>(arena + (size_t)n * kChunkBytes);
  T3D_CUDA(cudaMemcpyAsync(f, W[0], sizeof(float) * Sp::CIN * Sp::C1, cudaMemcpyDeviceToDevice, st));
  f += Sp::CIN * Sp::C1;
  T3D_CUDA(cudaMemcpyAsync(f, bias[0], sizeof(float) * Sp::C1, cudaMemcpyDeviceToDevice, st));
  f += Sp::C1;
  for (int l = 0; l < Sp::NH; ++l) {
    T3D_CUDA(cudaMemcpyAsync(f, bias[1 + l], sizeof(float) * Sp::HN(l), cudaMemcpyDeviceToDevice, st));
    f += Sp::HN(l);
  }
  T3D_CUDA(cudaMemcpyAsync(f, bias[1 + Sp::NH], sizeof(float) * Sp::FC, cudaMemcpyDeviceToDevice, st));
  return 0;
}

template <int KIND>
static int chain_launch(const ChainArgs& a, int out_elems, cudaStream_t st) {
  using Sp = ChainSpec<KIND>;
  using L = ChainSmem<Sp>;
  static int prepared[kMaxDevices] = {0};
  if (int e = once_per_device(prepared, [] {
        return (int)cudaFuncSetAttribute(chain_max_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL + 1024);
      }))
    return e;
  const int sms = xg_num_sms();
  T3D_CUDA(cudaMemsetAsync(a.out, 0, sizeof(float) * (size_t)out_elems, st));
  int grid = sms - (sms % kClusterSize);
  if (!a.tiles) {
    const int nt = a.B * ((a.N + L::TILE - 1) / L::TILE);
    const int need = ((nt + kClusterSize - 1) / kClusterSize) * kClusterSize;
    if (need < grid) grid = need;
  }
  chain_max_kernel<KIND><<<grid, kChainThreads, L::TOTAL + 1024, st>>>(a);
  T3D_CHECK_LAUNCH();
  return 0;
}

#define CHAIN_SWITCH(kind, EXPR)                         \
  switch (kind) {                                        \
    case CHAIN_SEG1: { constexpr int K_ = CHAIN_SEG1; EXPR; } break;   \
    case CHAIN_TNET: { constexpr int K_ = CHAIN_TNET; EXPR; } break;   \
    case CHAIN_BOX: { constexpr int K_ = CHAIN_BOX; EXPR; } break;     \
    case CHAIN_BOXPC: { constexpr int K_ = CHAIN_BOXPC; EXPR; } break; \
    case CHAIN_BOXPCB: { constexpr int K_ = CHAIN_BOXPCB; EXPR; } break; \
    default: return T3D_ERR_ARG;                         \
  }

extern "C" size_t t3d_chain_arena_bytes(int kind) {
  size_t r = 0;
  switch (kind) {
    case CHAIN_SEG1: r = chain_arena_bytes<ChainSpec<CHAIN_SEG1>>(); break;
    case CHAIN_TNET: r = chain_arena_bytes<ChainSpec<CHAIN_TNET>>(); break;
    case CHAIN_BOX: r = chain_arena_bytes<ChainSpec<CHAIN_BOX>>(); break;
    case CHAIN_BOXPC: r = chain_arena_bytes<ChainSpec<CHAIN_BOXPC>>(); break;
    case CHAIN_BOXPCB: r = chain_arena_bytes<ChainSpec<CHAIN_BOXPCB>>(); break;
    default: r = 0;
  }
  return r;
}
extern "C" int t3d_chain_num_layers(int kind) {
  switch (kind) {
    case CHAIN_SEG1: return 2 + ChainSpec<CHAIN_SEG1>::NH;
    case CHAIN_TNET: return 2 + ChainSpec<CHAIN_TNET>::NH;
    case CHAIN_BOX: return 2 + ChainSpec<CHAIN_BOX>::NH;
    case CHAIN_BOXPC: return 2 + ChainSpec<CHAIN_BOXPC>::NH;
    case CHAIN_BOXPCB: return 2 + ChainSpec<CHAIN_BOXPCB>::NH;
    default: return T3D_ERR_ARG;
  }
}
extern "C" int t3d_chain_tile_points(int kind) {
  switch (kind) {
    case CHAIN_SEG1: return 128 * ChainSpec<CHAIN_SEG1>::NSUB;
    case CHAIN_TNET: return 128 * ChainSpec<CHAIN_TNET>::NSUB;
    case CHAIN_BOX: return 128 * ChainSpec<CHAIN_BOX>::NSUB;
    case CHAIN_BOXPC: return 128 * ChainSpec<CHAIN_BOXPC>::NSUB;
    case CHAIN_BOXPCB: return 128 * ChainSpec<CHAIN_BOXPCB>::NSUB;
    default: return T3D_ERR_ARG;
  }
}
extern "C" int t3d_chain_out_channels(int kind) {
  switch (kind) {
    case CHAIN_SEG1: return ChainSpec<CHAIN_SEG1>::FC;
    case CHAIN_TNET: return ChainSpec<CHAIN_TNET>::FC;
    case CHAIN_BOX: return ChainSpec<CHAIN_BOX>::FC;
    case CHAIN_BOXPC: return ChainSpec<CHAIN_BOXPC>::FC;
    case CHAIN_BOXPCB: return ChainSpec<CHAIN_BOXPCB>::FC;
    default: return T3D_ERR_ARG;
  }
}

extern "C" int t3d_pack_chain(int kind, const float* const* W, const float* const* bias, void* arena, t3d_stream_t stream) {
  if (!W || !bias || !arena) return T3D_ERR_ARG;
  if ((uintptr_t)arena & 1023) return T3D_ERR_ALIGN;
  CHAIN_SWITCH(kind, return pack_chain_impl<K_>(W, bias, reinterpret_cast<uint8_t*>(arena), S(stream)));
  return 0;
}

static int chain_max_bf16_impl(int kind, const float* pc, const uint8_t* rgb, int B, int N, int C, const float* center, const int* idx,
                               int idx_stride, const int* count, const void* tiles, const int* num_tiles,
                               const float* box_center, const float* box_dims, const float* box_orient, const void* arena,
                               float* out, void* emit, t3d_stream_t stream);
extern "C" int t3d_chain_max_bf16(int kind, const float* pc, int B, int N, int C, const float* center, const int* idx,
                                  int idx_stride, const int* count, const void* tiles, const int* num_tiles,
                                  const float* box_center, const float* box_dims, const float* box_orient, const void* arena,
                                  float* out, void* emit, t3d_stream_t stream) {
  return chain_max_bf16_impl(kind, pc, nullptr, B, N, C, center, idx, idx_stride, count, tiles, num_tiles, box_center, box_dims, box_orient,
                             arena, out, emit, stream);
}
extern "C" int t3d_chain_max_bf16_wire(int kind, const float* xyz, const uint8_t* rgb, int B, int N, const float* center, const int* idx,
                                       int idx_stride, const int* count, const void* tiles, const int* num_tiles,
                                       const float* box_center, const float* box_dims, const float* box_orient, const void* arena,
                                       float* out, void* emit, t3d_stream_t stream) {
  if (!rgb) return T3D_ERR_ARG;
  return chain_max_bf16_impl(kind, xyz, rgb, B, N, 6, center, idx, idx_stride, count, tiles, num_tiles, box_center, box_dims, box_orient,
                             arena, out, emit, stream);
}
static int chain_max_bf16_impl(int kind, const float* pc, const uint8_t* rgb, int B, int N, int C, const float* center, const int* idx,
                               int idx_stride, const int* count, const void* tiles, const int* num_tiles,
                               const float* box_center, const float* box_dims, const float* box_orient, const void* arena,
                               float* out, void* emit, t3d_stream_t stream) {
  if (!pc || !arena || !out) return T3D_ERR_ARG;
  if (rgb && kind != CHAIN_SEG1 && kind != CHAIN_BOXPCB) return T3D_ERR_ARG;      // the chains whose raw input has 6 channels
  if (B <= 0 || N <= 0) return T3D_ERR_SHAPE;
  if ((uintptr_t)arena & 15) return T3D_ERR_ALIGN;
  if ((tiles != nullptr) != (num_tiles != nullptr)) return T3D_ERR_ARG;
  if (idx && idx_stride <= 0) return T3D_ERR_SHAPE;
  if (kind == CHAIN_BOXPC && (!box_center || !box_dims || !box_orient)) return T3D_ERR_ARG;
  if ((kind == CHAIN_BOXPC || kind == CHAIN_BOXPCB || kind == CHAIN_SEG1) ? C != 6 : C < 3) return T3D_ERR_SHAPE;
  if (emit && (kind != CHAIN_SEG1 || ((uintptr_t)emit & 15))) return T3D_ERR_ARG;
  ChainArgs a{pc, B, N, C, center, idx, idx_stride, count, reinterpret_cast<const int4*>(tiles), num_tiles,
              box_center, box_dims, box_orient, reinterpret_cast<const uint8_t*>(arena), out,
              reinterpret_cast<__nv_bfloat16*>(emit), g_trace, rgb};
  CHAIN_SWITCH(kind, return chain_launch<K_>(a, B * ChainSpec<K_>::FC, S(stream)));
  return 0;
}

extern "C" int t3d_normalize_pc(const float* pc, int B, int N, int C, int mode, float* out, t3d_stream_t stream) {
  if (!pc || !out) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0 || C < 3 || (mode != 0 && mode != 1)) return T3D_ERR_SHAPE;
  normalize_pc_kernel<<<B, 256, 0, S(stream)>>>(pc, N, C, mode, out);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" size_t t3d_seg2_arena_bytes(void) { return kSeg2ArenaBytes; }

extern "C" int t3d_pack_seg2(const float* W6p, const float* W7, const float* W8, const float* W9, const float* b7,
                             const float* b8, const float* b9, const float* W10, const float* b10, void* arena,
                             t3d_stream_t stream) {
  if (!W6p || !W7 || !W8 || !W9 || !b7 || !b8 || !b9 || !W10 || !b10 || !arena) return T3D_ERR_ARG;
  if ((uintptr_t)arena & 1023) return T3D_ERR_ALIGN;
  cudaStream_t st = S(stream);
  PackTable tab;
  int n = 0;
  auto c6 = [&](int nb) { tab.d[n++] = PackDesc{W6p, 512, 64, 0, nb * 128, 128}; };
  auto c7 = [&](int nb) {      // per K-block: rows 0-127 then rows 128-255 (one N=256 operand over two ring stages)
    for (int kb = 0; kb < 2; ++kb)
      for (int nh = 0; nh < 2; ++nh) tab.d[n++] = PackDesc{W7, 256, 512, nb * 128 + kb * 64, nh * 128, 128};
  };
  // arena order (seg_stage2_pair.cuh / seg_stage2_x2.cuh index it by chunk id): c6(0) c6(1) c7(0) c6(2) c6(3) c7(1) c7(2) c7(3) conv8 conv9
  c6(0); c6(1); c7(0); c6(2); c6(3); c7(1); c7(2); c7(3);
  for (int kb = 0; kb < 4; ++kb) tab.d[n++] = PackDesc{W8, 128, 256, kb * 64, 0, 128};
  for (int kb = 0; kb < 2; ++kb) tab.d[n++] = PackDesc{W9, 128, 128, kb * 64, 0, 128};
  if (n != kSeg2Chunks) return T3D_ERR_SHAPE;
  uint8_t* ar = reinterpret_cast<uint8_t*>(arena);
  pack_table_kernel<<<n, 256, 0, st>>>(tab, ar);
  T3D_CHECK_LAUNCH();
  float* f = reinterpret_cast<float*>(ar + (size_t)n * kChunkBytes);
  T3D_CUDA(cudaMemcpyAsync(f, b7, 4 * 256, cudaMemcpyDeviceToDevice, st));
  T3D_CUDA(cudaMemcpyAsync(f + 256, b8, 4 * 128, cudaMemcpyDeviceToDevice, st));
  T3D_CUDA(cudaMemcpyAsync(f + 384, b9, 4 * 128, cudaMemcpyDeviceToDevice, st));
  T3D_CUDA(cudaMemcpyAsync(f + 512, W10, 4 * 256, cudaMemcpyDeviceToDevice, st));
  T3D_CUDA(cudaMemcpyAsync(f + 768, b10, 4 * 2, cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" int t3d_seg_stage2_bf16(const void* point_feat, const float* gbias, const void* arena, float* logits, int B, int N,
                                   t3d_stream_t stream) {
  if (!point_feat || !gbias || !arena || !logits) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0) return T3D_ERR_SHAPE;
  if (((uintptr_t)point_feat & 15) || ((uintptr_t)gbias & 15) || ((uintptr_t)logits & 7) || ((uintptr_t)arena & 15))
    return T3D_ERR_ALIGN;
  static int prepared[kMaxDevices] = {0};
  if (int e = once_per_device(prepared, [] {
        return (int)cudaFuncSetAttribute(seg_stage2_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Seg2QSmem::TOTAL + 1024);
      }))
    return e;
  const int sms = xg_num_sms();
  Seg2Args a{reinterpret_cast<const __nv_bfloat16*>(point_feat), gbias, reinterpret_cast<const uint8_t*>(arena), logits, B, N, g_trace};
  const int nt = B * ((N + 127) / 128);
  int grid = sms - (sms % kClusterSize);
  const int need = ((nt + kClusterSize - 1) / kClusterSize) * kClusterSize;
  if (need < grid) grid = need;
  seg_stage2_pair_kernel<<<grid, kSeg2PThreads, Seg2QSmem::TOTAL + 1024, S(stream)>>>(a);
  T3D_CHECK_LAUNCH();
  return 0;
}

// ----------------------------------------------------------------------------- f16x2 split-precision chains
// First-order correction of the tensor core's round-toward-zero accumulation: every K=16 step that adds into a non-empty
// accumulator loses ~0.36 ulp of it on average, so the epilogue scale of a layer with n such steps carries (1 + c n).
// c is calibrated against the float64 oracle (tests/gpu_mask_exactness.py; emulated in tests/numerics_split_study.py).
static float g_x2_debias = 2.1e-8f;
extern "C" int t3d_set_x2_debias(float c) {
  if (!(c >= 0.0f && c < 1e-6f)) return T3D_ERR_ARG;
  g_x2_debias = c;
  return 0;
}
extern "C" float t3d_get_x2_debias(void) { return g_x2_debias; }
// epilogue scale of a layer whose accumulator holds (s * true value): kbn K-blocks issued in groups of `group`
static float x2_inv(float s, int kbn, int group) {
  const int neff = 12 * kbn - 8 * (group < kbn ? group : kbn);      // steps after the first group's small products
  return (1.0f / s) * (1.0f + g_x2_debias * (float)neff);
}
static bool is_pow2(float x) { int e; return x > 0.0f && std::frexp(x, &e) == 0.5f; }

template <int KIND>
static int pack_chain_x2_impl(const float* const* W, const float* const* bias, const float* wscale, uint8_t* arena, cudaStream_t st) {
  using Sp = ChainSpec<KIND>;
  PackTable tab;
  int n = 0;
  // chunk order = consumption order of chain_max_x2_kernel: per output block, per group of K-blocks: lo images, then hi images
  auto block = [&](const float* Wl, int ldw, int ktot, int row0, int nrows, float sc) {
    const int kbn = ktot / 64;
    for (int g0 = 0; g0 < kbn; g0 += kX2Group) {
      const int gn = kbn - g0 < kX2Group ? kbn - g0 : kX2Group;
      for (int part = 2; part >= 1; --part)
        for (int j = 0; j < gn; ++j)
          if (n < kPackMax) tab.d[n++] = PackDesc{Wl, ldw, ktot, (g0 + j) * 64, row0, nrows, part, sc};
    }
  };
  for (int l = 0; l <= Sp::NH; ++l)
    if (!is_pow2(wscale[l])) return T3D_ERR_ARG;
  for (int l = 0; l < Sp::NH; ++l)
    for (int nb = 0; nb < (Sp::HN(l) + 127) / 128; ++nb)
      block(W[1 + l], Sp::HN(l), Sp::HK(l), nb * 128, Sp::HN(l) - nb * 128 < 128 ? Sp::HN(l) - nb * 128 : 128, wscale[l]);
  for (int mt = 0; mt < Sp::FC / 128; ++mt) block(W[1 + Sp::NH], Sp::FC, Sp::FK, mt * 128, 128, wscale[Sp::NH]);
  if (n != x2_num_chunks<Sp>()) return T3D_ERR_SHAPE;
  pack_table_kernel<<<n, 256, 0, st>>>(tab, arena);
  T3D_CHECK_LAUNCH();
  float* f = reinterpret_cast<float*>(arena + (size_t)n * kChunkBytes);
  int e;
  if ((e = scale_copy(f, W[0], Sp::CIN * Sp::C1, kX2ActScale, st)) != 0) return e;
  f += Sp::CIN * Sp::C1;
  if ((e = scale_copy(f, bias[0], Sp::C1, kX2ActScale, st)) != 0) return e;
  f += Sp::C1;
  for (int l = 0; l < Sp::NH; ++l) {
    if ((e = scale_copy(f, bias[1 + l], Sp::HN(l), kX2ActScale, st)) != 0) return e;
    f += Sp::HN(l);
  }
  if ((e = scale_copy(f, bias[1 + Sp::NH], Sp::FC, 1.0f, st)) != 0) return e;
  f += Sp::FC;
  float inv[4] = {0, 0, 0, 0};
  for (int l = 0; l < Sp::NH; ++l) inv[l] = x2_inv(wscale[l], Sp::HK(l) / 64, kX2Group);
  inv[Sp::NH] = x2_inv(wscale[Sp::NH] * kX2ActScale, Sp::FK / 64, kX2Group);
  write4_kernel<<<1, 1, 0, st>>>(f, inv[0], inv[1], inv[2], inv[3]);
  T3D_CHECK_LAUNCH();
  return 0;
}

template <int KIND>
static int chain_launch_x2(const ChainArgs& a, int out_elems, cudaStream_t st) {
  using Sp = ChainSpec<KIND>;
  using L = X2Smem<Sp>;
  static int prepared[kMaxDevices] = {0};
  if (int e = once_per_device(prepared, [] {
        return (int)cudaFuncSetAttribute(chain_max_x2_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL + 1024);
      }))
    return e;
  const int sms = xg_num_sms();
  T3D_CUDA(cudaMemsetAsync(a.out, 0, sizeof(float) * (size_t)out_elems, st));
  int grid = sms - (sms % kClusterSize);
  if (!a.tiles) {
    const int nt = a.B * ((a.N + 127) / 128);
    const int need = ((nt + kClusterSize - 1) / kClusterSize) * kClusterSize;
    if (need < grid) grid = need;
  }
  chain_max_x2_kernel<KIND><<<grid, kChainThreads, L::TOTAL + 1024, st>>>(a);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" size_t t3d_chain_arena_bytes_x2(int kind) {
  switch (kind) {
    case CHAIN_SEG1: return x2_arena_bytes<ChainSpec<CHAIN_SEG1>>();
    case CHAIN_TNET: return x2_arena_bytes<ChainSpec<CHAIN_TNET>>();
    case CHAIN_BOX: return x2_arena_bytes<ChainSpec<CHAIN_BOX>>();
    case CHAIN_BOXPC: return x2_arena_bytes<ChainSpec<CHAIN_BOXPC>>();
    case CHAIN_BOXPCB: return x2_arena_bytes<ChainSpec<CHAIN_BOXPCB>>();
    default: return 0;
  }
}

extern "C" int t3d_pack_chain_x2(int kind, const float* const* W, const float* const* bias, const float* wscale, void* arena,
                                 t3d_stream_t stream) {
  if (!W || !bias || !wscale || !arena) return T3D_ERR_ARG;
  if ((uintptr_t)arena & 1023) return T3D_ERR_ALIGN;
  CHAIN_SWITCH(kind, return pack_chain_x2_impl<K_>(W, bias, wscale, reinterpret_cast<uint8_t*>(arena), S(stream)));
  return 0;
}

extern "C" int t3d_chain_max_x2(int kind, const float* pc, int B, int N, int C, const float* center, const int* idx,
                                int idx_stride, const int* count, const void* tiles, const int* num_tiles,
                                const float* box_center, const float* box_dims, const float* box_orient, const void* arena,
                                float* out, void* emit, t3d_stream_t stream) {
  if (!pc || !arena || !out) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0) return T3D_ERR_SHAPE;
  if ((uintptr_t)arena & 15) return T3D_ERR_ALIGN;
  if ((tiles != nullptr) != (num_tiles != nullptr)) return T3D_ERR_ARG;
  if (idx && idx_stride <= 0) return T3D_ERR_SHAPE;
  if (kind == CHAIN_BOXPC && (!box_center || !box_dims || !box_orient)) return T3D_ERR_ARG;
  if ((kind == CHAIN_BOXPC || kind == CHAIN_BOXPCB || kind == CHAIN_SEG1) ? C != 6 : C < 3) return T3D_ERR_SHAPE;
  if (emit && (kind != CHAIN_SEG1 || ((uintptr_t)emit & 15))) return T3D_ERR_ARG;
  ChainArgs a{pc, B, N, C, center, idx, idx_stride, count, reinterpret_cast<const int4*>(tiles), num_tiles,
              box_center, box_dims, box_orient, reinterpret_cast<const uint8_t*>(arena), out,
              reinterpret_cast<__nv_bfloat16*>(emit), g_trace};
  CHAIN_SWITCH(kind, return chain_launch_x2<K_>(a, B * ChainSpec<K_>::FC, S(stream)));
  return 0;
}

extern "C" size_t t3d_seg2_arena_bytes_x2(void) { return kSeg2XArenaBytes; }

extern "C" int t3d_pack_seg2_x2(const float* W6p, const float* W7, const float* W8, const float* W9, const float* b7,
                                const float* b8, const float* b9, const float* W10, const float* b10, const float* wscale,
                                void* arena, t3d_stream_t stream) {
  if (!W6p || !W7 || !W8 || !W9 || !b7 || !b8 || !b9 || !W10 || !b10 || !wscale || !arena) return T3D_ERR_ARG;
  if ((uintptr_t)arena & 1023) return T3D_ERR_ALIGN;
  for (int l = 0; l < 4; ++l)
    if (!is_pow2(wscale[l])) return T3D_ERR_ARG;
  cudaStream_t st = S(stream);
  PackTable tab;
  int n = 0;
  auto c6 = [&](int nb) {
    for (int part = 2; part >= 1; --part) tab.d[n++] = PackDesc{W6p, 512, 64, 0, nb * 128, 128, part, wscale[0]};
  };
  auto c7 = [&](int kbg) {      // [lo rows 0-127][lo rows 128-255][hi rows 0-127][hi rows 128-255]
    for (int part = 2; part >= 1; --part)
      for (int nh = 0; nh < 2; ++nh) tab.d[n++] = PackDesc{W7, 256, 512, kbg * 64, nh * 128, 128, part, wscale[1]};
  };
  // consumption order of seg_stage2_x2_kernel
  c6(0); c6(1); c7(0); c7(1); c6(2); c7(2); c7(3); c6(3); c7(4); c7(5); c7(6); c7(7);
  for (int kb = 0; kb < 4; ++kb)
    for (int part = 2; part >= 1; --part) tab.d[n++] = PackDesc{W8, 128, 256, kb * 64, 0, 128, part, wscale[2]};
  for (int kb = 0; kb < 2; ++kb)
    for (int part = 2; part >= 1; --part) tab.d[n++] = PackDesc{W9, 128, 128, kb * 64, 0, 128, part, wscale[3]};
  if (n != kSeg2XChunks) return T3D_ERR_SHAPE;
  uint8_t* ar = reinterpret_cast<uint8_t*>(arena);
  pack_table_kernel<<<n, 256, 0, st>>>(tab, ar);
  T3D_CHECK_LAUNCH();
  float* f = reinterpret_cast<float*>(ar + (size_t)n * kChunkBytes);
  int e;
  if ((e = scale_copy(f, b7, 256, kX2ActScale, st)) != 0) return e;
  if ((e = scale_copy(f + 256, b8, 128, kX2ActScale, st)) != 0) return e;
  if ((e = scale_copy(f + 384, b9, 128, kX2ActScale, st)) != 0) return e;
  if ((e = scale_copy(f + 512, W10, 256, 1.0f / kX2ActScale, st)) != 0) return e;
  if ((e = scale_copy(f + 768, b10, 2, 1.0f, st)) != 0) return e;
  write4_kernel<<<1, 1, 0, st>>>(f + 770, x2_inv(wscale[0], 1, 1), x2_inv(wscale[1], 8, 1), x2_inv(wscale[2], 4, 1),
                                 x2_inv(wscale[3], 2, 1));
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_seg_stage2_x2(const void* point_feat, const float* gbias, const void* arena, float* logits, int B, int N,
                                 t3d_stream_t stream) {
  if (!point_feat || !gbias || !arena || !logits) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0) return T3D_ERR_SHAPE;
  if (((uintptr_t)point_feat & 15) || ((uintptr_t)gbias & 15) || ((uintptr_t)logits & 7) || ((uintptr_t)arena & 15))
    return T3D_ERR_ALIGN;
  static int prepared[kMaxDevices] = {0};
  if (int e = once_per_device(prepared, [] {
        return (int)cudaFuncSetAttribute(seg_stage2_x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Seg2XSmem::TOTAL + 1024);
      }))
    return e;
  const int sms = xg_num_sms();
  Seg2XArgs a{reinterpret_cast<const uint8_t*>(point_feat), gbias, reinterpret_cast<const uint8_t*>(arena), logits, B, N, g_trace};
  const int nt = B * ((N + 127) / 128);
  int grid = sms - (sms % kClusterSize);
  const int need = ((nt + kClusterSize - 1) / kClusterSize) * kClusterSize;
  if (need < grid) grid = need;
  seg_stage2_x2_kernel<<<grid, kSeg2XThreads, Seg2XSmem::TOTAL + 1024, S(stream)>>>(a);
  T3D_CHECK_LAUNCH();
  return 0;
}

// ----------------------------------------------------------------------------- training-step kernels
// Extras of the training-mode layers (t3d_gemm_bn_f32): lazy batch norm of the A operand and fused output statistics.
struct GemmBnExtras {
  const float* a_scale = nullptr; const float* a_shift = nullptr;      // A := relu(a_scale[c] * A + a_shift[c])
  float* st_sum = nullptr; float* st_sq = nullptr; const float* st_shift = nullptr;
  unsigned* pool_max = nullptr; unsigned* pool_min = nullptr; int pool_rows = 0;      // fused max-pool keys (C may be null)
  bool any() const { return a_scale != nullptr || st_sum != nullptr; }
};

static int gemm_f32_impl(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C,
                         int ldc, int M, int N, int K, int splitk, const float* bias, void* ws, size_t ws_bytes,
                         t3d_stream_t stream, const GemmBnExtras& x) {
  if (!A || !B || (!C && !x.pool_max)) return T3D_ERR_ARG;
  if (M <= 0 || N <= 0 || K <= 0 || splitk <= 0 || ldc < N) return T3D_ERR_SHAPE;
  if ((sam != 1 && sak != 1) || (sbk != 1 && sbn != 1)) return T3D_ERR_SHAPE;
  if (splitk > 1) T3D_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * ldc, S(stream)));
  GemmArgs a{A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, splitk, bias, x.st_sum, x.st_sq, x.st_shift, x.pool_max, x.pool_min, x.pool_rows};
  if (x.st_sum != nullptr) {
    T3D_CUDA(cudaMemsetAsync(x.st_sum, 0, sizeof(float) * N, S(stream)));
    T3D_CUDA(cudaMemsetAsync(x.st_sq, 0, sizeof(float) * N, S(stream)));
  }
  if (x.pool_max != nullptr) {
    // the fused max-pool lives in the statistics epilogue of the tensor-core path: whole tiles inside one group
    if (!x.st_sum || !x.pool_min || x.pool_rows <= 0 || x.pool_rows % kXgBM != 0 || M % x.pool_rows != 0 || splitk != 1 || sak != 1 ||
        !xg_fits(M, N, K) || (K <= kSkinnyMax))
      return T3D_ERR_SHAPE;
    const size_t groups = (size_t)(M / x.pool_rows);
    T3D_CUDA(cudaMemsetAsync(x.pool_max, 0x00, sizeof(unsigned) * groups * N, S(stream)));
    T3D_CUDA(cudaMemsetAsync(x.pool_min, 0xff, sizeof(unsigned) * groups * N, S(stream)));
  }
  {   // HBM-bound first-layer shapes (skinny_gemm.cuh); 128-bit accesses need aligned bases and leading dimensions
    auto al16 = [](const void* p) { return (((uintptr_t)p) & 15) == 0; };
    const int sms = xg_num_sms();
    if (M >= 4096 && K <= kSkinnyMax && sak == 1 && sbn == 1 && N % 4 == 0 && N <= 1024 && 256 % (N / 4) == 0 && al16(C) && ldc % 4 == 0 &&
        sizeof(float) * (size_t)(K * N + 3 * N) <= kSkinnySmemMax && x.a_scale == nullptr) {
      if (splitk > 1) { /* C was zeroed above; a single pass writes every element */ }
      skinny_k_kernel<<<sms * 8, 256, sizeof(float) * (size_t)(K * N + 3 * N), S(stream)>>>(A, sam, B, sbk, bias, C, ldc, M, N, K, 0,
                                                                                             x.st_sum, x.st_sq, x.st_shift);
      T3D_CHECK_LAUNCH();
      return 0;
    }
    if (x.any() && !xg_fits(M, N, K)) return T3D_ERR_SHAPE;      // the extras exist on the tensor-core path (and skinny_k) only
    // (either layout of the small operand: dgrad passes W as [N, K] rows, a forward layer with <= 16 outputs -- inst_seg conv10,
    // 128 -> 2 -- as [K, N]; the kernel stages it in shared memory by element strides)
    if (!x.any() && M >= 4096 && N <= kSkinnyMax && sak == 1 && K % 4 == 0 && K <= 1024 && al16(A) && sam % 4 == 0 && splitk == 1 &&
        sizeof(float) * (size_t)N * K <= kSkinnySmemMax) {
      skinny_n_kernel<<<sms * 8, 256, sizeof(float) * (size_t)N * K, S(stream)>>>(A, sam, B, sbn, bias, C, ldc, M, N, K, sbk);
      T3D_CHECK_LAUNCH();
      return 0;
    }
    if (!x.any() && K >= 4096 && M <= kSkinnyMax && sam == 1 && sbn == 1 && !bias && N % 4 == 0 && N <= 1024 && 256 % (N / 4) == 0 && al16(B) &&
        sbk % 4 == 0 && (size_t)M * N * sizeof(float) <= kSkinnySmemMax) {
      if (splitk <= 1) T3D_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * ldc, S(stream)));
      skinny_m_kernel<<<sms * 4, 256, sizeof(float) * (size_t)M * N, S(stream)>>>(A, sak, B, sbk, C, ldc, M, N, K);
      T3D_CHECK_LAUNCH();
      return 0;
    }
  }
  if (xg_fits(M, N, K)) {                    // tensor cores, bf16 x 3 split (xgemm.cuh)
    static int prepared[kMaxDevices] = {0};
    if (int e = once_per_device(prepared, [] {
          return xg_prepare(xgemm_kernel<true, true, 3>) | xg_prepare(xgemm_kernel<true, false, 3>) |
                 xg_prepare(xgemm_kernel<false, true, 3>) | xg_prepare(xgemm_kernel<false, false, 3>) |
                 xg_prepare(xgemm_kernel<true, true, 2>) | xg_prepare(xgemm_kernel<true, false, 2>) |
                 xg_prepare(xgemm_kernel<false, true, 2>) | xg_prepare(xgemm_kernel<false, false, 2>) |
                 xg_prepare(xgemm_kernel<true, true, 1>) | xg_prepare(xgemm_kernel<true, false, 1>) |
                 xg_prepare(xgemm_kernel<false, true, 1>) | xg_prepare(xgemm_kernel<false, false, 1>) |
                 xg_prepare(xgemm_pre_kernel<3>) | xg_prepare(xgemm_pre_kernel<2>) | xg_prepare(xgemm_pre_kernel<1>) |
                 xg_prepare_pp<3>(xg_as_kernel<3, false>) | xg_prepare_pp<2>(xg_as_kernel<2, false>) | xg_prepare_pp<1>(xg_as_kernel<1, false>) |
                 xg_prepare_pp<3>(xg_pp_kernel<3, false>) | xg_prepare_pp<2>(xg_pp_kernel<2, false>) | xg_prepare_pp<1>(xg_pp_kernel<1, false>);
        }))
      return e;
    // The tensor core adds each 16-deep partial sum into the fp32 accumulator with truncation, a bias of ~0.5 ulp per
    // step (measured: 700 ulp of sum|a||b| after K = 20000, 25 after K = 600); K chunks are kept <= 2048 so that long
    // reductions (wgrad over B*N rows) are summed across chunks by round-to-nearest fp32 reductions instead.
    int sk = splitk;
    if ((K + sk - 1) / sk > kXgMaxKChunk) sk = (K + kXgMaxKChunk - 1) / kXgMaxKChunk;
    if (sk > 1 && splitk == 1) T3D_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * ldc, S(stream)));
    a.splitk = sk;
    const int kchunk = ((K + sk - 1) / sk + kXgBK - 1) / kXgBK * kXgBK;
    const int nz = (K + kchunk - 1) / kchunk;            // every K chunk non-empty
    const bool ak = (sak == 1), bk = (sbk == 1 && sbn != 1);
    const int ntm = (M + kXgBM - 1) / kXgBM, ntn = (N + kXgBN - 1) / kXgBN;
    const long long lda = ak ? sam : sak, ldb = bk ? sbn : sbk;
    XgOperands o{A, lda, B, ldb, M, N, K, kchunk, (ak && xg_aligned16(A, lda)) ? 1 : 0, (bk && xg_aligned16(B, ldb)) ? 1 : 0, ntn, g_trace, nullptr, 0,
                 x.a_scale, x.a_shift};
    const dim3 grid((unsigned)ntm * ntn, nz);
    // CTA-pair kernel (256 x 256 tiles, xgemm.cuh) for the engines with at most two operand pieces: every shape with at
    // least 256 rows and columns except the K <= 128 forward / dgrad layers, which stay on the persistent kernels,
    // and the problems with fewer pair CTAs than SMs (FC heads: a handful of 256 x 256 tiles each walking the whole K range alone)
    const bool pair = xg_parts() <= 2 && M >= 256 && N >= 256 && !(ak && nz == 1 && sk == 1 && K <= 128 && xg_pre_ok(M, N, K, ws, ws_bytes)) &&
                      (long long)((M + 255) / 256) * ((N + 255) / 256) * 2 * nz >= xg_num_sms();
    const bool pre = !pair && ak && nz == 1 && sk == 1 && xg_pre_ok(M, N, K, ws, ws_bytes);
    if (pair) {
      if (x.a_scale && K % kXgBK != 0) return T3D_ERR_SHAPE;
      if (x.st_sum && sk != 1) return T3D_ERR_SHAPE;
      static int prepared_pair[kMaxDevices] = {0};
      if (int e = once_per_device(prepared_pair, [] {
            auto prep = [](auto kern) { return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kXgSmemBytes + 24u * 1024u)); };
            return prep(xgemm_pair_kernel<true, true, 2>) | prep(xgemm_pair_kernel<true, false, 2>) |
                   prep(xgemm_pair_kernel<false, true, 2>) | prep(xgemm_pair_kernel<false, false, 2>) |
                   prep(xgemm_pair_kernel<true, true, 1>) | prep(xgemm_pair_kernel<true, false, 1>) |
                   prep(xgemm_pair_kernel<false, true, 1>) | prep(xgemm_pair_kernel<false, false, 1>);
          }))
        return e;
#ifndef XG_PAIR_SOLO
#define XG_PAIR_SOLO 0
#endif
      // XG_PAIR_SOLO: request more than half of the shared memory so that only one cluster is resident per SM pair
      constexpr uint32_t kPairSmem = kXgSmemBytes + (XG_PAIR_SOLO ? 24u * 1024u : 0u);
      const int ntm2 = (M + 255) / 256, ntn2 = (N + 255) / 256;
      o.ntn = ntn2;
      const dim3 pgrid((unsigned)ntm2 * ntn2 * 2, nz);
#define XG_LAUNCH_PAIR(P)                                                                                     \
  do {                                                                                                        \
    if (ak && bk) xgemm_pair_kernel<true, true, P><<<pgrid, kXgThreads, kPairSmem, S(stream)>>>(a, o);      \
    else if (ak) xgemm_pair_kernel<true, false, P><<<pgrid, kXgThreads, kPairSmem, S(stream)>>>(a, o);      \
    else if (bk) xgemm_pair_kernel<false, true, P><<<pgrid, kXgThreads, kPairSmem, S(stream)>>>(a, o);      \
    else xgemm_pair_kernel<false, false, P><<<pgrid, kXgThreads, kPairSmem, S(stream)>>>(a, o);             \
  } while (0)
      if (xg_parts() == 2) XG_LAUNCH_PAIR(2);
      else XG_LAUNCH_PAIR(1);
#undef XG_LAUNCH_PAIR
      T3D_CHECK_LAUNCH();
      return 0;
    }
    if (x.any()) {
      // lazy BN of a k-contiguous A lives in the pre-split-B loaders (128-bit parameter loads: K % 32 == 0, aligned arrays);
      // of a row-contiguous A (wgrad) in the generic loader.  Statistics need the whole K range in one CTA.
      if (x.a_scale && ak && (!pre || K % kXgBK != 0 || K > kXgLazyMaxK)) return T3D_ERR_SHAPE;
      if (x.a_scale && !ak && pre) return T3D_ERR_SHAPE;
      if (x.st_sum && sk != 1) return T3D_ERR_SHAPE;
      if (x.pool_max && !pre) return T3D_ERR_SHAPE;
    }
    if (pre) {     // forward / dgrad: pre-split B
      o.bpre = reinterpret_cast<const uint8_t*>(ws);
      o.nkb = (K + 63) / 64;
      xg_presplit_kernel<<<dim3(o.nkb, ntn, 4), 256, 0, S(stream)>>>(B, ldb, bk ? 1 : 0, N, K, xg_parts(), reinterpret_cast<uint8_t*>(ws));
      if (xg_use_as(K, ntn)) {
        const int g = ntm < xg_num_sms() ? ntm : xg_num_sms();
        XG_BY_PARTS((xg_as_kernel<P, false><<<g, XgPP<P>::THREADS, XgPP<P>::SMEM, S(stream)>>>(a, LinearArgs{}, o)));
      } else if (xg_use_pp(K)) {
        const int g = ntm * ntn < xg_num_sms() ? ntm * ntn : xg_num_sms();
        XG_BY_PARTS((xg_pp_kernel<P, false><<<g, XgPP<P>::THREADS, XgPP<P>::SMEM, S(stream)>>>(a, LinearArgs{}, o)));
      } else XG_BY_PARTS((xgemm_pre_kernel<P><<<grid, kXgThreads, kXgSmemBytes, S(stream)>>>(a, o)));
      T3D_CHECK_LAUNCH();
      return 0;
    }
#define XG_LAUNCH(P)                                                                                     \
  do {                                                                                                   \
    if (ak && bk) xgemm_kernel<true, true, P><<<grid, kXgThreads, kXgSmemBytes, S(stream)>>>(a, o);        \
    else if (ak) xgemm_kernel<true, false, P><<<grid, kXgThreads, kXgSmemBytes, S(stream)>>>(a, o);        \
    else if (bk) xgemm_kernel<false, true, P><<<grid, kXgThreads, kXgSmemBytes, S(stream)>>>(a, o);        \
    else xgemm_kernel<false, false, P><<<grid, kXgThreads, kXgSmemBytes, S(stream)>>>(a, o);               \
  } while (0)
    XG_BY_PARTS(XG_LAUNCH(P));
#undef XG_LAUNCH
    T3D_CHECK_LAUNCH();
    return 0;
  }
  if (M >= 128 && N >= 96 && K >= 16) {      // 128 x 128 tiles (sgemm.cuh)
    const int kchunk = ((K + splitk - 1) / splitk + kSgBK - 1) / kSgBK * kSgBK;
    const bool ak = (sak == 1), bk = (sbk == 1 && sbn != 1);
    SgemmOperands o{A, sam, sak, B, sbk, sbn, M, N, K, kchunk, sg_aligned16(A, ak ? sam : sak) ? 1 : 0, sg_aligned16(B, bk ? sbn : sbk) ? 1 : 0};
    dim3 grid((M + kSgBM - 1) / kSgBM, (N + kSgBN - 1) / kSgBN, splitk);
    if (ak && bk) gemm128_kernel<true, true><<<grid, 256, 0, S(stream)>>>(a, o);
    else if (ak) gemm128_kernel<true, false><<<grid, 256, 0, S(stream)>>>(a, o);
    else if (bk) gemm128_kernel<false, true><<<grid, 256, 0, S(stream)>>>(a, o);
    else gemm128_kernel<false, false><<<grid, 256, 0, S(stream)>>>(a, o);
    T3D_CHECK_LAUNCH();
    return 0;
  }
  dim3 grid((M + 63) / 64, (N + 63) / 64, splitk);
  gemm_f32_kernel<<<grid, 256, 0, S(stream)>>>(a);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_gemm_f32_ws(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C,
                               int ldc, int M, int N, int K, int splitk, const float* bias, void* ws, size_t ws_bytes,
                               t3d_stream_t stream) {
  return gemm_f32_impl(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, splitk, bias, ws, ws_bytes, stream, GemmBnExtras{});
}

// Can t3d_gemm_bn_f32 serve this problem?  kind 0: forward (k-contiguous A, lazy BN by k and / or output statistics),
// kind 1: wgrad (row-contiguous A, lazy BN by row).  Mirrors the dispatch of gemm_f32_impl.
extern "C" int t3d_gemm_bn_supported(int M, int N, int K, int kind) {
  if (kind == 0) {
    if (M >= 4096 && K <= kSkinnyMax && N % 4 == 0 && N <= 1024 && 256 % (N / 4) == 0 &&
        sizeof(float) * (size_t)(K * N + 3 * N) <= kSkinnySmemMax)
      return 2;                                     // first-layer kernel: statistics only (its input is never a lazy BN)
    return (xg_fits(M, N, K) && M >= 4096 && K <= kXgLazyMaxK && K % kXgBK == 0) ? 1 : 0;
  }
  if (kind == 1) return (xg_fits(M, N, K) && !(K >= 4096 && M <= kSkinnyMax)) ? 1 : 0;
  return 0;
}

extern "C" int t3d_gemm_bn_f32(const float* A, long long sam, long long sak, const float* a_scale, const float* a_shift,
                               const float* B, long long sbk, long long sbn, float* C, int ldc, int M, int N, int K, int splitk,
                               const float* bias, float* st_sum, float* st_sq, const float* st_shift, void* ws, size_t ws_bytes,
                               t3d_stream_t stream) {
  if ((a_scale != nullptr) != (a_shift != nullptr) || (st_sum != nullptr) != (st_sq != nullptr)) return T3D_ERR_ARG;
  GemmBnExtras x;
  x.a_scale = a_scale; x.a_shift = a_shift; x.st_sum = st_sum; x.st_sq = st_sq; x.st_shift = st_shift;
  return gemm_f32_impl(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, splitk, bias, ws, ws_bytes, stream, x);
}

// Forward GEMM of a forward-only lazy BN layer whose output only feeds the max-pool over the pool_rows rows of each group
// (the segmentation net's conv5 in the semi-supervised step): statistics AND the per-group column max / min come out of the
// epilogue, the [M, N] output is never written.  t3d_pool_bn_finish turns the keys into the pooled activations.
extern "C" int t3d_gemm_bn_pool_f32(const float* A, long long lda, const float* a_scale, const float* a_shift, const float* W, int ldw,
                                    int M, int N, int K, const float* bias, float* st_sum, float* st_sq, const float* st_shift,
                                    int pool_rows, void* pool_max_keys, void* pool_min_keys, void* ws, size_t ws_bytes,
                                    t3d_stream_t stream) {
  if ((a_scale != nullptr) != (a_shift != nullptr) || !st_sum || !st_sq || !pool_max_keys || !pool_min_keys) return T3D_ERR_ARG;
  GemmBnExtras x;
  x.a_scale = a_scale; x.a_shift = a_shift; x.st_sum = st_sum; x.st_sq = st_sq; x.st_shift = st_shift;
  x.pool_max = reinterpret_cast<unsigned*>(pool_max_keys); x.pool_min = reinterpret_cast<unsigned*>(pool_min_keys); x.pool_rows = pool_rows;
  return gemm_f32_impl(A, lda, 1, W, ldw, 1, nullptr, N, M, N, K, 1, bias, ws, ws_bytes, stream, x);
}
extern "C" int t3d_pool_bn_finish(const void* pool_max_keys, const void* pool_min_keys, const float* a_scale, const float* a_shift,
                                  int groups, int C, float* out, t3d_stream_t stream) {
  if (!pool_max_keys || !pool_min_keys || !a_scale || !a_shift || !out) return T3D_ERR_ARG;
  if (groups <= 0 || C <= 0) return T3D_ERR_SHAPE;
  pool_bn_finish_kernel<<<(groups * C + 255) / 256, 256, 0, S(stream)>>>(reinterpret_cast<const unsigned*>(pool_max_keys),
                                                                        reinterpret_cast<const unsigned*>(pool_min_keys), a_scale, a_shift,
                                                                        groups * C, C, out);
  T3D_CHECK_LAUNCH();
  return 0;
}

// y0[n] = sum_k a(k) W[k, n] + bias[n] for ONE row a (lazy BN applied when a_scale != null): the shift of the fused statistics
extern "C" int t3d_row0(const float* a, const float* a_scale, const float* a_shift, const float* W, int ldw, const float* bias,
                        int K, int N, float* y0, t3d_stream_t stream) {
  if (!a || !W || !y0 || K <= 0 || N <= 0) return T3D_ERR_ARG;
  row0_kernel<<<(N + 31) / 32, 256, 0, S(stream)>>>(a, a_scale, a_shift, W, ldw, bias, K, N, y0);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_gemm_f32(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C,
                            int ldc, int M, int N, int K, int splitk, const float* bias, t3d_stream_t stream) {
  return t3d_gemm_f32_ws(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, splitk, bias, nullptr, 0, stream);
}

static int colstats_impl(const float* X, const float* out, const float* y, const float* mean, const float* rstd, float* o0,
                         float* o1, int M, int C, int mode, int act, const float* a_scale, const float* a_shift, t3d_stream_t stream) {
  if (!X || !o0 || !o1 || (mode == 1 && (!y || !mean || !rstd))) return T3D_ERR_ARG;
  if (M <= 0 || C <= 0 || (mode != 0 && mode != 1) || act < 0 || act > 3) return T3D_ERR_SHAPE;
  if ((a_scale != nullptr) != (a_shift != nullptr) || (a_scale && (out || mode != 1 || act != 1))) return T3D_ERR_ARG;
  T3D_CUDA(cudaMemsetAsync(o0, 0, sizeof(float) * C, S(stream)));
  T3D_CUDA(cudaMemsetAsync(o1, 0, sizeof(float) * C, S(stream)));
  ColStatArgs a{X, out, y, mean, rstd, o0, o1, M, C, mode, act, a_scale, a_shift};
  int chunks = (M + 511) / 512;
  if (chunks > 1024) chunks = 1024;
  dim3 grid((C + 31) / 32, chunks);
  colstats_kernel<<<grid, 256, 0, S(stream)>>>(a);
  T3D_CHECK_LAUNCH();
  return 0;
}
extern "C" int t3d_colstats(const float* X, const float* out, const float* y, const float* mean, const float* rstd, float* o0,
                            float* o1, int M, int C, int mode, int act, t3d_stream_t stream) {
  return colstats_impl(X, out, y, mean, rstd, o0, o1, M, C, mode, act, nullptr, nullptr, stream);
}
// BN-backward reductions of a lazy BN layer (no stored output): ReLU mask = a_scale * y + a_shift > 0
extern "C" int t3d_colstats_lazy(const float* dOut, const float* y, const float* mean, const float* rstd, const float* a_scale,
                                 const float* a_shift, float* s1, float* s2, int M, int C, t3d_stream_t stream) {
  if (!a_scale || !a_shift) return T3D_ERR_ARG;
  return colstats_impl(dOut, nullptr, y, mean, rstd, s1, s2, M, C, 1, 1, a_scale, a_shift, stream);
}

extern "C" int t3d_bn_finalize(const float* sum, const float* sumsq, const float* shift, int M, int C, float eps, float decay,
                               float* mean, float* rstd, float* moving_mean, float* moving_var, t3d_stream_t stream) {
  if (!sum || !sumsq || !mean || !rstd || ((moving_mean != nullptr) != (moving_var != nullptr))) return T3D_ERR_ARG;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, S(stream)>>>(sum, sumsq, shift, M, C, eps, decay, mean, rstd, moving_mean, moving_var);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_bn_apply(const float* y, const float* mean, const float* rstd, const float* gamma, const float* beta,
                            float* out, int M, int C, int act, t3d_stream_t stream) {
  if (!y || !mean || !rstd || !gamma || !beta || !out) return T3D_ERR_ARG;
  if (act < 0 || act > 3) return T3D_ERR_SHAPE;
  const size_t total = (size_t)M * C;
  if (ew4_ok(total, C, {y, out}) && C <= 2048) {
    using F4 = const float4*;
    bn_apply4_kernel<<<(ew4_grid(total) + kBnIter - 1) / kBnIter, 256, sizeof(float) * 4 * (size_t)C, S(stream)>>>((F4)y, mean, rstd, gamma, beta, (float4*)out,
                                                             (unsigned)(total / 4), (unsigned)(C / 4), act);
  } else {
    bn_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(y, mean, rstd, gamma, beta, out, total, C, act);
  }
  T3D_CHECK_LAUNCH();
  return 0;
}

static int bn_backward_impl(float* dOut, const float* out, const float* y, const float* mean, const float* rstd,
                            const float* gamma, const float* s1, const float* s2, int M, int C, int act, const float* a_scale,
                            const float* a_shift, t3d_stream_t stream) {
  if (!dOut || !y || !mean || !rstd || !gamma || !s1 || !s2) return T3D_ERR_ARG;
  if (act < 0 || act > 3) return T3D_ERR_SHAPE;
  if ((a_scale != nullptr) != (a_shift != nullptr) || (a_scale && (out || act != 1))) return T3D_ERR_ARG;
  const size_t total = (size_t)M * C;
  if (ew4_ok(total, C, {dOut, out, y}) && C <= (a_scale ? 1536 : 2048)) {
    using F4 = const float4*;
    bn_backward4_kernel<<<(ew4_grid(total) + kBnIter - 1) / kBnIter, 256, sizeof(float) * (a_scale ? 7 : 5) * (size_t)C, S(stream)>>>((float4*)dOut, (F4)out, (F4)y, mean, rstd, gamma, s1, s2,
                                                                (unsigned)(total / 4), (unsigned)(C / 4), M, act, a_scale, a_shift);
  } else {
    bn_backward_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(dOut, out, y, mean, rstd, gamma, s1, s2, total, C, M, act, a_scale, a_shift);
  }
  T3D_CHECK_LAUNCH();
  return 0;
}
extern "C" int t3d_bn_backward(float* dOut, const float* out, const float* y, const float* mean, const float* rstd,
                               const float* gamma, const float* s1, const float* s2, int M, int C, int act, t3d_stream_t stream) {
  return bn_backward_impl(dOut, out, y, mean, rstd, gamma, s1, s2, M, C, act, nullptr, nullptr, stream);
}
extern "C" int t3d_bn_backward_lazy(float* dOut, const float* y, const float* mean, const float* rstd, const float* gamma,
                                    const float* a_scale, const float* a_shift, const float* s1, const float* s2, int M, int C,
                                    t3d_stream_t stream) {
  if (!a_scale || !a_shift) return T3D_ERR_ARG;
  return bn_backward_impl(dOut, nullptr, y, mean, rstd, gamma, s1, s2, M, C, 1, a_scale, a_shift, stream);
}

extern "C" int t3d_bn_finalize_affine(const float* sum, const float* sumsq, const float* shift, int M, int C, float eps, float decay,
                                      const float* gamma, const float* beta, float* mean, float* rstd, float* a_scale,
                                      float* a_shift, float* moving_mean, float* moving_var, t3d_stream_t stream) {
  if (!sum || !sumsq || !gamma || !beta || !mean || !rstd || !a_scale || !a_shift || ((moving_mean != nullptr) != (moving_var != nullptr)))
    return T3D_ERR_ARG;
  bn_finalize_affine_kernel<<<(C + 127) / 128, 128, 0, S(stream)>>>(sum, sumsq, shift, M, C, eps, decay, gamma, beta, mean, rstd, a_scale,
                                                                  a_shift, moving_mean, moving_var);
  T3D_CHECK_LAUNCH();
  return 0;
}

// Backward of [BN -> ReLU -> (x rowmask) -> max over the N rows of each group] of a lazy BN layer (train_ops.cuh): writes the
// BN input gradient dY [B*N, C] and s1 = d beta, s2 = d gamma from the pooled gradient g [B, C] and the arg-max rows.
extern "C" int t3d_pool_bn_backward(const float* g, const int* arg, const float* rowmask, const float* y, const float* mean,
                                    const float* rstd, const float* gamma, const float* a_scale, const float* a_shift, int B, int N,
                                    int C, float* s1, float* s2, float* dY, t3d_stream_t stream) {
  if (!g || !arg || !y || !mean || !rstd || !gamma || !a_scale || !a_shift || !s1 || !s2 || !dY) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0 || C <= 0) return T3D_ERR_SHAPE;
  T3D_CUDA(cudaMemsetAsync(s1, 0, sizeof(float) * C, S(stream)));
  T3D_CUDA(cudaMemsetAsync(s2, 0, sizeof(float) * C, S(stream)));
  PoolBnBwdArgs a{g, arg, rowmask, y, mean, rstd, gamma, a_scale, a_shift, s1, s2, dY, B, N, C};
  int chunks = (B + 7) / 8;
  if (chunks > 64) chunks = 64;
  pool_bn_stats_kernel<<<dim3((C + 31) / 32, chunks), dim3(32, 8), 0, S(stream)>>>(a);
  const int M = B * N;
  const size_t total = (size_t)M * C;
  if (ew4_ok(total, C, {y, dY}) && C <= 2048) {
    const unsigned n4 = (unsigned)(total / 4);
    pool_bn_dense_kernel<<<(n4 + 4095) / 4096, 256, sizeof(float) * 3 * (size_t)C, S(stream)>>>((const float4*)y, mean, rstd, gamma, s1, s2,
                                                                                             (float4*)dY, n4, (unsigned)(C / 4), M);
  } else {
    pool_bn_dense_scalar_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(y, mean, rstd, gamma, s1, s2, dY, total, C, M);
  }
  pool_bn_scatter_kernel<<<(B * C + 255) / 256, 256, 0, S(stream)>>>(a);
  T3D_CHECK_LAUNCH();
  return 0;
}

// `keys`: caller-owned scratch of B * C 64-bit words for the row-split kernel (null: the serial kernel)
static int maxpool_fwd_impl(const float* x, const float* rowmask, int B, int N, int C, float* out, int* arg, const float* a_scale,
                            const float* a_shift, void* keys, t3d_stream_t stream) {
  const int cblocks = (C + 255) / 256;
  int split = 1;
  if (keys != nullptr && (((uintptr_t)keys) & 7) == 0) {
    split = (xg_num_sms() * 8 + cblocks * B - 1) / (cblocks * B);
    if (split > N / 64) split = N / 64;
  }
  if (split <= 1) {
    maxpool_fwd_kernel<<<(B * C + 255) / 256, 256, 0, S(stream)>>>(x, rowmask, B, N, C, out, arg, a_scale, a_shift);
    T3D_CHECK_LAUNCH();
    return 0;
  }
  const int rows = (N + split - 1) / split;
  T3D_CUDA(cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * (size_t)B * C, S(stream)));
  maxpool_split_kernel<<<dim3(cblocks, B, (N + rows - 1) / rows), 256, 0, S(stream)>>>(x, rowmask, B, N, C, rows,
                                                                                     reinterpret_cast<unsigned long long*>(keys), a_scale, a_shift);
  maxpool_split_finish_kernel<<<(B * C + 255) / 256, 256, 0, S(stream)>>>(reinterpret_cast<const unsigned long long*>(keys), B * C, out, arg);
  T3D_CHECK_LAUNCH();
  return 0;
}
extern "C" int t3d_maxpool_masked_fwd(const float* x, const float* rowmask, int B, int N, int C, float* out, int* arg,
                                      t3d_stream_t stream) {
  if (!x || !out || !arg) return T3D_ERR_ARG;
  return maxpool_fwd_impl(x, rowmask, B, N, C, out, arg, nullptr, nullptr, nullptr, stream);
}
// the same with a caller-owned scratch of 8 * B * C bytes: rows are split across blocks when B * C alone cannot fill the GPU;
// a_scale / a_shift != NULL: x is the pre-BN tensor of a lazy BN layer (pooled = max relu(a_scale x + a_shift) * rowmask)
extern "C" int t3d_maxpool_fwd_ws(const float* x, const float* a_scale, const float* a_shift, const float* rowmask, int B, int N, int C,
                                  float* out, int* arg, void* keys, t3d_stream_t stream) {
  if (!x || !out || !arg || (a_scale != nullptr) != (a_shift != nullptr)) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0 || C <= 0) return T3D_ERR_SHAPE;
  return maxpool_fwd_impl(x, rowmask, B, N, C, out, arg, a_scale, a_shift, keys, stream);
}
// pooled = max_n relu(a_scale[c] * y + a_shift[c]) (* rowmask): the max-pool reads the pre-BN tensor of a lazy BN layer
extern "C" int t3d_maxpool_lazy_fwd(const float* y, const float* a_scale, const float* a_shift, const float* rowmask, int B, int N,
                                    int C, float* out, int* arg, t3d_stream_t stream) {
  if (!y || !a_scale || !a_shift || !out || !arg) return T3D_ERR_ARG;
  maxpool_fwd_kernel<<<(B * C + 255) / 256, 256, 0, S(stream)>>>(y, rowmask, B, N, C, out, arg, a_scale, a_shift);
  T3D_CHECK_LAUNCH();
  return 0;
}
extern "C" int t3d_maxpool_fwd(const float* x, int B, int N, int C, float* out, int* arg, t3d_stream_t stream) {
  return t3d_maxpool_masked_fwd(x, nullptr, B, N, C, out, arg, stream);
}

extern "C" int t3d_maxpool_masked_bwd(const float* dout, const int* arg, const float* rowmask, int B, int N, int C, float* dx,
                                      t3d_stream_t stream) {
  if (!dout || !arg || !dx) return T3D_ERR_ARG;
  T3D_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * N * C, S(stream)));
  maxpool_bwd_kernel<<<(B * C + 255) / 256, 256, 0, S(stream)>>>(dout, arg, rowmask, B, N, C, dx);
  T3D_CHECK_LAUNCH();
  return 0;
}
extern "C" int t3d_maxpool_bwd(const float* dout, const int* arg, int B, int N, int C, float* dx, t3d_stream_t stream) {
  return t3d_maxpool_masked_bwd(dout, arg, nullptr, B, N, C, dx, stream);
}

extern "C" int t3d_pool_rows(const int* arg, int B, int N, int C, int nslot, int* rows, int* slot, int* count, t3d_stream_t stream) {
  if (!arg || !rows || !slot || !count) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0 || C <= 0 || nslot < (C < N ? C : N)) return T3D_ERR_SHAPE;
  const int words = (N + 31) / 32;
  if (sizeof(unsigned) * 2 * (size_t)words > 48 * 1024) return T3D_ERR_SHAPE;
  pool_rows_kernel<<<B, 256, sizeof(unsigned) * 2 * words, S(stream)>>>(arg, N, C, nslot, rows, slot, count);
  T3D_CHECK_LAUNCH();
  return 0;
}
extern "C" int t3d_gather_rows(const float* src, const int* rows, int B, int N, int nslot, int C, float* dst, t3d_stream_t stream) {
  if (!src || !rows || !dst) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0 || nslot <= 0 || C <= 0) return T3D_ERR_SHAPE;
  if (C % 4 == 0 && ((((uintptr_t)src) | ((uintptr_t)dst)) & 15)) return T3D_ERR_ALIGN;
  const size_t total = (size_t)B * nslot, items = total * (size_t)(C % 4 == 0 ? C / 4 : C);
  gather_rows_kernel<<<(unsigned)((items + 255) / 256), 256, 0, S(stream)>>>(src, rows, N, nslot, C, total, dst);
  T3D_CHECK_LAUNCH();
  return 0;
}
extern "C" int t3d_scatter_pool_grad(const float* g, const int* slot, int B, int C, int nslot, float* dst, t3d_stream_t stream) {
  if (!g || !slot || !dst) return T3D_ERR_ARG;
  if (B <= 0 || C <= 0 || nslot <= 0) return T3D_ERR_SHAPE;
  T3D_CUDA(cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)B * nslot * C, S(stream)));
  scatter_pool_grad_kernel<<<(B * C + 255) / 256, 256, 0, S(stream)>>>(g, slot, B, C, nslot, dst);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_scale_mask(const float* x, const float* mask, float scale, float* out, long long n, t3d_stream_t stream) {
  if (!x || !mask || !out || n <= 0) return T3D_ERR_ARG;
  if (ew4_ok((size_t)n, 4, {x, mask, out})) {
    scale_mask4_kernel<<<ew4_grid((size_t)n), 256, 0, S(stream)>>>((const float4*)x, (const float4*)mask, scale, (float4*)out, (unsigned)(n / 4));
  } else {
    scale_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, S(stream)>>>(x, mask, scale, out, (size_t)n);
  }
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_boxpc_loss(const t3d_boxpc_loss_args* p, t3d_stream_t stream) {
  if (!p || !p->out9 || !p->y_iou || !p->y_dc || !p->y_ds || !p->y_da || !p->total) return T3D_ERR_ARG;
  if (p->B <= 0) return T3D_ERR_SHAPE;
  T3D_CUDA(cudaMemsetAsync(p->total, 0, sizeof(float), S(stream)));
  BoxpcLossArgs a{p->out9, p->y_iou, p->y_dc, p->y_ds, p->y_da, p->B, p->fit_bound, p->w_cls, p->w_delta, p->wc, p->ws, p->wa,
                  p->huber, p->cls_losses, p->delta_losses, p->total, p->grad, p->pred_weigh, p->loss_weigh, p->stop_grad};
  if (p->pred_weigh < 0 || p->pred_weigh > 1 || p->loss_weigh < 0 || p->loss_weigh > 2) return T3D_ERR_ARG;
  boxpc_loss_kernel<<<(p->B + 127) / 128, 128, 0, S(stream)>>>(a);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_adam(float* param, const float* grad, float* m, float* v, long long n, float lr_t, float beta1, float beta2,
                        float eps, float grad_scale, t3d_stream_t stream) {
  if (!param || !grad || !m || !v || n <= 0) return T3D_ERR_ARG;
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, S(stream)>>>(param, grad, m, v, (size_t)n, lr_t, beta1, beta2, eps, grad_scale);
  T3D_CHECK_LAUNCH();
  return 0;
}

// ----------------------------------------------------------------------------- semi-supervised step: losses and helpers
extern "C" int t3d_seg_ce(const float* logits, const int* labels, int B, int N, float* out, t3d_stream_t stream) {
  if (!logits || !labels || !out) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0) return T3D_ERR_SHAPE;
  if ((uintptr_t)logits & 7) return T3D_ERR_ALIGN;
  seg_ce_kernel<<<B, 256, 0, S(stream)>>>(logits, labels, N, out);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_class_dims_stats(const float* dims_reg, const float* one_hot, int B, int NC, float* cls_sum, float* cls_cnt,
                                    t3d_stream_t stream) {
  if (!dims_reg || !one_hot || !cls_sum || !cls_cnt) return T3D_ERR_ARG;
  if (B <= 0 || NC <= 0 || NC > 32) return T3D_ERR_SHAPE;
  T3D_CUDA(cudaMemsetAsync(cls_sum, 0, sizeof(float) * 3 * NC, S(stream)));
  T3D_CUDA(cudaMemsetAsync(cls_cnt, 0, sizeof(float) * NC, S(stream)));
  class_dims_stats_kernel<<<(B + 127) / 128, 128, 0, S(stream)>>>(dims_reg, one_hot, B, NC, cls_sum, cls_cnt);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_semi_loss(const t3d_semi_loss_args* p, t3d_stream_t stream) {
  if (!p || !p->out || !p->stage1_center || !p->one_hot || !p->y_center || !p->y_orient_cls || !p->y_orient_reg || !p->y_dims_cls ||
      !p->y_dims_reg || !p->is_data_2D || !p->mean_size || !p->dF || !p->ds1 || !p->g_reg || !p->total)
    return T3D_ERR_ARG;
  if (p->w_reproj != 0.f && (!p->Rtilt || !p->K || !p->rot_frust || !p->box2D || !p->img_dim)) return T3D_ERR_ARG;
  if (p->B <= 0 || p->NH <= 0 || p->NH > kMaxNH || p->NS <= 0 || p->NS > kMaxNS || p->NC <= 0 || p->NC > 32) return T3D_ERR_SHAPE;
  T3D_CUDA(cudaMemsetAsync(p->total, 0, sizeof(float) * 8, S(stream)));
  SemiLossArgs a;
  a.out = p->out; a.stage1_center = p->stage1_center; a.mask_losses = p->mask_losses; a.one_hot = p->one_hot;
  a.y_center = p->y_center; a.y_orient_cls = p->y_orient_cls; a.y_orient_reg = p->y_orient_reg; a.y_dims_cls = p->y_dims_cls;
  a.y_dims_reg = p->y_dims_reg; a.Rtilt = p->Rtilt; a.K = p->K; a.rot_frust = p->rot_frust; a.box2D = p->box2D; a.img_dim = p->img_dim;
  a.is_data_2D = p->is_data_2D; a.fit_logits = p->fit_logits; a.mean_size = p->mean_size; a.cls_sum = p->cls_sum; a.cls_cnt = p->cls_cnt; a.reg_in = p->reg_in;
  a.B = p->B; a.NH = p->NH; a.NS = p->NS; a.NC = p->NC; a.icv_train_mask = p->icv_train_mask;
  a.w_ce = p->w_ce; a.box_mult = p->box_mult; a.w_center = p->w_center; a.w_ocls = p->w_ocls; a.w_dcls = p->w_dcls; a.w_oreg = p->w_oreg;
  a.w_dreg = p->w_dreg; a.w_tnet = p->w_tnet; a.w_corner = p->w_corner; a.weak_mult = p->weak_mult; a.w_icv = p->w_icv;
  a.w_reproj = p->w_reproj; a.w_fit = p->w_fit; a.reproj_only_2d = p->reproj_only_2d; a.fit_only_2d = p->fit_only_2d;
  a.use_softmax_proj = p->use_softmax_proj; a.softmax_scale = p->softmax_scale; a.dilate = p->dilate; a.clip_lower_b = p->clip_lower_b;
  a.clip_pred_box = p->clip_pred_box; a.reproj_mse = p->reproj_mse; a.icv_mse = p->icv_mse; a.train_box_mask = p->train_box_mask;
  a.inv_n3d = p->inv_n3d; a.dF = p->dF; a.ds1 = p->ds1; a.g_reg = p->g_reg; a.dfit = p->dfit; a.per_sample = p->per_sample; a.total = p->total;
  semi_loss_kernel<<<(p->B + 63) / 64, 64, 0, S(stream)>>>(a);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_box_reg_backward(const float* out, const float* g_reg, const float* mean_size, int B, int NH, int NS, float* dF,
                                    float* ds1, t3d_stream_t stream) {
  if (!out || !g_reg || !mean_size || !dF) return T3D_ERR_ARG;
  if (B <= 0 || NH <= 0 || NS <= 0) return T3D_ERR_SHAPE;
  box_reg_backward_kernel<<<(B + 127) / 128, 128, 0, S(stream)>>>(out, g_reg, mean_size, B, NH, NS, dF, ds1);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_boxpc_features_bwd(const float* pc, int B, int N, int C, const float* center, const float* orient, const float* g6,
                                      float* g_box, t3d_stream_t stream) {
  if (!pc || !center || !orient || !g6 || !g_box) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0 || C < 3) return T3D_ERR_SHAPE;
  boxpc_features_bwd_kernel<<<B, 256, 0, S(stream)>>>(pc, N, C, center, orient, g6, g_box);
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_act_bwd(float* dout, const float* out, long long n, int act, t3d_stream_t stream) {
  if (!dout || !out || n <= 0) return T3D_ERR_ARG;
  if (act < 0 || act > 3) return T3D_ERR_SHAPE;
  if (ew4_ok((size_t)n, 4, {dout, out})) {
    act_bwd4_kernel<<<ew4_grid((size_t)n), 256, 0, S(stream)>>>((float4*)dout, (const float4*)out, (unsigned)(n / 4), act);
  } else {
    act_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, S(stream)>>>(dout, out, (size_t)n, act);
  }
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_rowmask_mul(const float* x, const float* rowmask, float* out, long long M, int C, t3d_stream_t stream) {
  if (!x || !rowmask || !out || M <= 0 || C <= 0) return T3D_ERR_ARG;
  const size_t total = (size_t)M * C;
  if (ew4_ok(total, C, {x, out})) {
    rowmask_mul4_kernel<<<ew4_grid(total), 256, 0, S(stream)>>>((const float4*)x, rowmask, (float4*)out, (unsigned)(total / 4), (unsigned)(C / 4));
  } else {
    rowmask_mul_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(x, rowmask, out, total, C);
  }
  T3D_CHECK_LAUNCH();
  return 0;
}

extern "C" int t3d_soft_mask(const float* logits, int B, int N, float* out, t3d_stream_t stream) {
  if (!logits || !out) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0 || (((uintptr_t)logits) & 7)) return T3D_ERR_SHAPE;
  const size_t n = (size_t)B * N;
  soft_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, S(stream)>>>(logits, n, out);
  T3D_CHECK_LAUNCH();
  return 0;
}
extern "C" int t3d_seg_ce_bwd(const float* logits, const int* labels, const float* w, const float* gmask, int B, int N,
                              float* dlogits, t3d_stream_t stream) {
  if (!logits || !labels || !w || !dlogits) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0 || (((uintptr_t)logits | (uintptr_t)dlogits) & 7)) return T3D_ERR_SHAPE;
  const size_t n = (size_t)B * N;
  seg_ce_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, S(stream)>>>(logits, labels, w, gmask, B, N, dlogits);
  T3D_CHECK_LAUNCH();
  return 0;
}
extern "C" int t3d_group_colsum(const float* x, int B, int N, int C, float* out, t3d_stream_t stream) {
  if (!x || !out) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0 || C <= 0) return T3D_ERR_SHAPE;
  T3D_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * C, S(stream)));
  const int cb = (C + 255) / 256;
  int split = (xg_num_sms() * 8 + cb * B - 1) / (cb * B);
  if (split > N / 32) split = N / 32;
  if (split < 1) split = 1;
  const int rows = (N + split - 1) / split;
  group_colsum_kernel<<<dim3(cb, B, (N + rows - 1) / rows), 256, 0, S(stream)>>>(x, N, C, rows, out);
  T3D_CHECK_LAUNCH();
  return 0;
}
extern "C" int t3d_group_sum(const float* x, int B, int N, int C, float scale, float* out, t3d_stream_t stream) {
  if (!x || !out) return T3D_ERR_ARG;
  if (B <= 0 || N <= 0 || C <= 0 || C > 8) return T3D_ERR_SHAPE;
  group_sum_kernel<<<B, 256, 0, S(stream)>>>(x, N, C, scale, out);
  T3D_CHECK_LAUNCH();
  return 0;
}
