/* t3d_b200 -- C ABI of the B200-native Frustum-PointNet hot path of yewsiang/Transferable3D.
 *
 * The reference has no FFI: its boundary is Python functions that build a TF1 graph plus
 * sess.run (sunrgbd_detection/test_semisup.py:221-226, train_boxpc.py:365-366).  Each entry point
 * below replaces the TF ops behind one of those reference functions (cited per function) and is
 * what a binding for this path would bind.
 *
 * Conventions: every pointer is a DEVICE pointer to contiguous memory unless marked `host`;
 * the caller owns all buffers including workspaces (the library never allocates); `stream` is a
 * cudaStream_t (NULL = legacy default stream); calls are asynchronous on that stream and
 * re-entrant for distinct streams/buffers.  Return value: 0 ok, <0 invalid argument
 * (T3D_ERR_*), >0 a cudaError_t.  float = IEEE fp32, int = int32.
 */
#ifndef T3D_B200_H_
#define T3D_B200_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define T3D_VERSION 100
#define T3D_ERR_ARG (-1)      /* null pointer / bad enum */
#define T3D_ERR_SHAPE (-2)    /* unsupported shape */
#define T3D_ERR_ALIGN (-3)    /* pointer not aligned as required */

typedef void* t3d_stream_t;

int t3d_version(void);
const char* t3d_error_string(int code);

/* activation ids of tf_util.conv2d / fully_connected activation_fn (models/tf_util.py:1321,1497) */
enum { T3D_ACT_NONE = 0, T3D_ACT_RELU = 1, T3D_ACT_LEAKY_RELU = 2, T3D_ACT_TANH = 3 };

/* One conv2d(1x1) / fully_connected layer in fp32 with BN folded by the caller
 * (models/tf_util.py:1258-1323, 1463-1499):  Y = act(X.W + bias + gbias[row / rows_per_group]) * rowmask[row].
 * W is [K,N] row-major (the TF [Cin,Cout] layout).  If gmax != NULL (zero-initialised [groups,N]) the
 * max over the rows of each group is accumulated there (tf_util.max_pool2d over the points,
 * models/tf_util.py:1501-1524, on post-ReLU / masked values >= 0); Y may then be NULL. */
int t3d_linear_f32(const float* X, int ldx, const float* W, int ldw, const float* bias,
                   const float* gbias, int rows_per_group, float* Y, int ldy, int M, int K, int N,
                   int act, const float* rowmask, float* gmax, t3d_stream_t stream);

/* semisup_models.subtract_points_mean (semisup_models.py:145-162) and model_util.point_cloud_masking
 * (model_util.py:241-272): mask = float(l0 < l1), count, mean = sum(mask*xyz)/max(count,1),
 * xyz_stage1 = xyz - mean, plus idx = ascending indices of the masked-in points (compaction).
 * Any of mask / count / mean / xyz_stage1 / idx may be NULL. */
int t3d_mask_centroid(const float* logits, const float* pc, int B, int N, int C, float* mask,
                      int* count, float* mean, float* xyz_stage1, int* idx, t3d_stream_t stream);

/* model_util.tf_gather_object_pc / mask_to_indices (model_util.py:61-91).
 * mode 0: Philox4x32-10 counter RNG keyed (seed, frustum) -- see oracle/model_util.py;
 * mode 1: `choice` holds rank-space choice arrays [B,npoints] drawn by the host (numpy stream).
 * indices [B,npoints,2] = {frustum, point}; object_pc [B,npoints,c_out] = (xyz - mean, features). */
int t3d_resample(const int* idx, const int* count, int B, int N, int npoints, int mode, uint64_t seed,
                 const int* choice, int* indices, const float* pc, int C, const float* mean, int c_out,
                 float* object_pc, t3d_stream_t stream);

/* tile table for the ragged tcgen05 chains: tiles = int32[<= sum ceil(count/tile_pts)][4] */
int t3d_build_tiles(const int* count, int B, int tile_pts, void* tiles, int* num_tiles, t3d_stream_t stream);

/* fp32-mode inputs: xyz - center (semisup_models.py:204-209) and the BoxPC representation
 * tf_util.tf_get_box_pc_representation (tf_util.py:764-795): out[B,N,C+6]. */
int t3d_prepare_xyz(const float* pc, int B, int N, int C, const float* center, float* out, t3d_stream_t stream);
int t3d_boxpc_features(const float* pc, int B, int N, int C, const float* center, const float* dims,
                       const float* orient, float* out, t3d_stream_t stream);

/* Output parsing (semisup_models.py:265-290, semisup_v1_sunrgbd.py:203-222, model_util.parse_output_to_tensors
 * model_util.py:178-210) fused with tf_convert_box_params_from_anchor_to_reg_format_multi (tf_util.py:1001-1041). */
typedef struct {
  const float* output;         /* [B, 3+2NH+4NS] */
  const float* stage1_center;  /* [B,3] or NULL */
  const float* mean_size;      /* [NS,3] */
  const float* orient_anchors; /* [NH] */
  int B, NH, NS;
  float *center, *heading_scores, *heading_res_norm, *heading_res, *size_scores, *size_res_norm, *size_res;
  float *reg_center, *reg_dims, *reg_orient;
} t3d_parse_args;
int t3d_parse_box(const t3d_parse_args* args /* host */, t3d_stream_t stream);

int t3d_anchor_to_reg(const float* center, const float* dims_cls, const float* dims_reg, const float* orient_cls,
                      const float* orient_reg, const float* dims_anchors, const float* orient_anchors, int B, int NS,
                      int NH, float* out_center, float* out_dims, float* out_orient, t3d_stream_t stream);

/* boxpc_sunrgbd.get_model output slicing (boxpc_sunrgbd.py:70-95) + one refine step of
 * test_semisup.get_model (test_semisup.py:116-134). */
typedef struct {
  const float* out9;
  int B, weigh_pred_by_conf, weigh_during_test;
  float *fit_logits, *fit_prob; int* pred_fit;
  float *delta_center, *delta_size, *delta_angle;
  float *box_center, *box_dims, *box_orient;
  float *tot_center, *tot_size, *tot_angle;
} t3d_refine_args;
int t3d_boxpc_refine(const t3d_refine_args* args /* host */, t3d_stream_t stream);
/* F2_* end points (test_semisup.py:136-142) */
int t3d_f2(const float* f_center, const float* f_hres, const float* f_sres, const float* tot_center,
           const float* tot_angle, const float* tot_size, int B, int NH, int NS, float* f2_center,
           float* f2_hres, float* f2_sres, t3d_stream_t stream);

/* model_util.get_box3d_corners_helper (model_util.py:94-119) and get_box3d_corners(_sunrgbd) (:121-167) */
int t3d_box3d_corners_helper(const float* centers, const float* headings, const float* sizes, int n, float* out,
                             t3d_stream_t stream);
int t3d_box3d_corners_all(const float* center, const float* heading_res, const float* size_res, const float* mean_size,
                          const float* orient_anchors, int B, int NH, int NS, float* out, t3d_stream_t stream);

/* ---- oriented 3D box IoU and the BoxPC perturbation sampler ------------------------------------------------------
 * The reference calls box_util.box3d_iou (roi_seg_box3d_dataset.py:15,133; box_pc_fit_dataset.py:17,38-42), a module that is
 * absent from its tree: it is train/box_util.py of charlesq34/frustum-pointnets (Sutherland-Hodgman clip of the
 * bird's-eye-view rectangles, polygon area, overlap of the y extents); restated in oracle/box_util.py.
 * get_3d_box (roi_seg_box3d_dataset.py:84-100): size (l,w,h), heading about y, center -> corners [B,8,3]. */
int t3d_get_3d_box(const float* size, const float* heading, const float* center, int B, float* corners, t3d_stream_t stream);
/* box3d_iou on corner sets [B,8,3] x [B,8,3] -> iou3d [B], iou2d [B] (either may be NULL) */
int t3d_box3d_iou(const float* corners1, const float* corners2, int B, float* iou3d, float* iou2d, t3d_stream_t stream);
/* roi_seg_box3d_dataset.compute_box3d_iou (:102-139), the py_func behind semisup_v1_sunrgbd.get_iou_summary (:232-246):
 * argmax-select, class2angle / class2size, get_3d_box on prediction and label, box3d_iou -> iou2ds, iou3ds [B] */
typedef struct {
  const float *center_pred, *heading_logits, *heading_residuals, *size_logits, *size_residuals;
  const float* center_label; const int* heading_class_label; const float* heading_residual_label;
  const int* size_class_label; const float* size_residual_label;
  const float* mean_size;      /* [NS,3] */
  int B, NH, NS;
  float *iou2ds, *iou3ds;
} t3d_compute_iou_args;
int t3d_compute_box3d_iou(const t3d_compute_iou_args* args /* host */, t3d_stream_t stream);
/* BoxPCFitDataset.perturb_box_to_diff_ious (box_pc_fit_dataset.py:211-244): per box, draw (center, size, angle) perturbations
 * scaled by 1 - mean(band) until the 3D IoU with the original lies strictly inside bounds[b] = (lo, hi).  Randomness:
 * Philox4x32-10 keyed by seed with counter (attempt, 0|1, box, 0) -- see oracle/box_pc_fit_dataset.py.  attempts[b] = number
 * of draws used, or -1 if max_attempts draws all missed the band (the reference would loop for ever). */
typedef struct {
  const float *center, *size, *heading, *bounds;    /* [B,3], [B,3], [B], [B,2] */
  int B, max_attempts;
  float center_perturbation, size_perturbation, angle_perturbation;
  uint64_t seed;
  float *new_center, *new_size, *new_heading, *iou3d, *d_center, *d_size, *d_angle;
  int* attempts;
} t3d_perturb_args;
int t3d_perturb_boxes(const t3d_perturb_args* args /* host */, t3d_stream_t stream);

/* ---- tcgen05 (bf16 in / fp32 accumulate) fused per-point MLP chains ------------------------------------
 * kind: 0 inst_seg conv1-5 + max (semisup_models.py:76-97), 1 tnet convs + max (:172-189, model_util.py:300-316),
 *       2 box_est convs + max (:224-245), 3 box_pc_mask_model convs + max with the BoxPC features computed
 *       in the first-layer prologue (:334-376).
 * Weights are packed once into a caller-owned arena (BN already folded; W[l] is [K,N] row-major fp32 on device). */
enum { T3D_CHAIN_SEG1 = 0, T3D_CHAIN_TNET = 1, T3D_CHAIN_BOX = 2, T3D_CHAIN_BOXPC = 3,
       T3D_CHAIN_BOXPCB = 4 /* BoxPC representation B: conv 6-128-128-256-512 + max on the raw points (semisup_models.py:423-443) */ };
size_t t3d_chain_arena_bytes(int kind);
int t3d_chain_num_layers(int kind);     /* layer 1 + hidden + final */
int t3d_chain_tile_points(int kind);
int t3d_chain_out_channels(int kind);
int t3d_pack_chain(int kind, const float* const* W /* host array of device ptrs */,
                   const float* const* bias /* host array of device ptrs */, void* arena, t3d_stream_t stream);
/* out [B,FC] is zeroed by the call. idx/count/tiles/num_tiles describe compacted (or gathered) points;
 * all NULL = dense over the N points.  emit (seg1 only): bf16 [B*N,64] point_feat for stage 2. */
int t3d_chain_max_bf16(int kind, const float* pc, int B, int N, int C, const float* center, const int* idx,
                       int idx_stride, const int* count, const void* tiles, const int* num_tiles,
                       const float* box_center, const float* box_dims, const float* box_orient,
                       const void* arena, float* out, void* emit, t3d_stream_t stream);
/* Same with the 6-channel input in the wire format of t3d_assemble_points (xyz [B,N,3] fp32 + rgb [B,N,3] uint8, colour = k / 255
 * by IEEE division: bit-identical to assembling the (B,N,6) placeholder of semisup_v1_sunrgbd.py:39 first): the first layer of
 * the inst_seg chain (kind 0) / BoxPC representation B (kind 4) reads it directly, so the fp32 (B,N,6) tensor is never built. */
int t3d_chain_max_bf16_wire(int kind, const float* xyz, const uint8_t* rgb, int B, int N, const float* center,
                            const int* idx, int idx_stride, const int* count, const void* tiles, const int* num_tiles,
                            const float* box_center, const float* box_dims, const float* box_orient,
                            const void* arena, float* out, void* emit, t3d_stream_t stream);

/* inst_seg conv6'..conv10 (semisup_models.py:107-135) with conv6's global half folded into gbias [B,512]. */
size_t t3d_seg2_arena_bytes(void);
int t3d_pack_seg2(const float* W6p /* [64,512] */, const float* W7 /* [512,256] */, const float* W8 /* [256,128] */,
                  const float* W9 /* [128,128] */, const float* b7, const float* b8, const float* b9,
                  const float* W10 /* [128,2] */, const float* b10, void* arena, t3d_stream_t stream);
int t3d_seg_stage2_bf16(const void* point_feat /* bf16 [B*N,64] */, const float* gbias, const void* arena,
                        float* logits, int B, int N, t3d_stream_t stream);

/* ---- split-precision ("f16x2") variants of the same fused chains: fp32-accurate on the tensor cores --------------
 * Same reference functions, arguments and outputs as t3d_chain_max_bf16 / t3d_seg_stage2_bf16.  Every operand is a pair
 * of fp16 images x = hi + lo (22+ significant bits); a layer issues three tcgen05 products (hi.lo, lo.hi, hi.hi) into one
 * fp32 accumulator (csrc/chain_x2.cuh, csrc/seg_stage2_x2.cuh).  Tile = 128 points for every kind.
 * wscale (HOST array of powers of two, one per tensor-core layer: hidden..., final) scales the weights into the fp16
 * normal range; the arena stores the inverse.  emit (seg1) = per 128-point tile [hi image 16 KB][lo image 16 KB] of
 * conv3's output (32 KB x B x ceil(N/128)); gbias of the stage-2 call must be multiplied by 16 (the activation scale:
 * pass 16*W6[64:], 16*b6 to t3d_linear_f32).
 * t3d_set_x2_debias: per-accumulation-step correction of the tensor core's round-toward-zero accumulation folded into
 * the packed epilogue scales (applies to arenas packed afterwards; default 2.1e-8, 0 = off). */
size_t t3d_chain_arena_bytes_x2(int kind);
int t3d_pack_chain_x2(int kind, const float* const* W /* host array of device ptrs */,
                      const float* const* bias /* host array of device ptrs */, const float* wscale /* host */,
                      void* arena, t3d_stream_t stream);
int t3d_chain_max_x2(int kind, const float* pc, int B, int N, int C, const float* center, const int* idx,
                     int idx_stride, const int* count, const void* tiles, const int* num_tiles,
                     const float* box_center, const float* box_dims, const float* box_orient,
                     const void* arena, float* out, void* emit, t3d_stream_t stream);
size_t t3d_seg2_arena_bytes_x2(void);
int t3d_pack_seg2_x2(const float* W6p, const float* W7, const float* W8, const float* W9, const float* b7, const float* b8,
                     const float* b9, const float* W10, const float* b10, const float* wscale /* host[4]: conv6' 7 8 9 */,
                     void* arena, t3d_stream_t stream);
int t3d_seg_stage2_x2(const void* point_feat, const float* gbias /* x16 */, const void* arena, float* logits, int B, int N,
                      t3d_stream_t stream);
int t3d_set_x2_debias(float per_step);
float t3d_get_x2_debias(void);

/* ---- training-step kernels (fp32, CUDA cores) ---------------------------------------------------------
 * The TF ops behind train_boxpc.train (train_boxpc.py:219-300) and the trainable part of train_semisup_adv.train:
 * conv2d / fully_connected forward, dgrad and wgrad as one strided GEMM, training-mode batch norm
 * (tf.contrib.layers.batch_norm, tf_util.py:1645-1664) fused with ReLU, max_pool2d with its gradient, dropout
 * (tf_util.py:1720-1741), the BoxPC loss (boxpc_sunrgbd.py:106-193) and tf.train.AdamOptimizer. */
/* C[M,N] (+)= sum_k A(m,k) B(k,n), element strides (sam,sak) / (sbk,sbn) with one of each pair == 1;
 * splitk > 1 accumulates partial sums atomically (C is zeroed by the call). */
int t3d_gemm_f32(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C,
                 int ldc, int M, int N, int K, int splitk, const float* bias, t3d_stream_t stream);
/* Engine behind t3d_linear_f32 / t3d_gemm_f32 for problems with M >= 128, N >= 64, K >= 32 (process-wide setting):
 *   1 (default) = tcgen05 tensor cores with every fp32 operand split exactly into three bf16 pieces and the six leading
 *                 partial products accumulated in fp32 (csrc/xgemm.cuh): fp32-level accuracy at 1/6 of the bf16 rate;
 *   0           = CUDA-core SGEMM (csrc/sgemm.cuh);
 *   2           = tcgen05 with the operands rounded once to bf16 (round to nearest) and fp32 accumulation: the bf16 mode
 *                 of the training steps (rel 1e-2 class accuracy), one tensor-core pass instead of six;
 *   3           = tcgen05 with every operand split into two bf16 pieces, both rounded to nearest (x = h + m + O(2^-18 x)),
 *                 and the three leading partial products: results within ~1e-5 of the fp32 value relative to
 *                 sum_k |a||b| (30 x tighter than TF32) at half the tensor-core work of engine 1.
 * The reference computes these layers with tf.nn.conv2d / tf.matmul in fp32 (models/tf_util.py:1308,1489). */
int t3d_set_f32_engine(int engine);
/* Same calls with a caller-owned device workspace (16-byte aligned, >= t3d_gemm_ws_bytes(N, K) bytes, private to the stream):
 * when A / X is k-contiguous, M >= 4096 and K <= 2048 the tensor-core engines split the small operand once per call into
 * the workspace instead of once per row tile.  ws == NULL is the plain call. */
size_t t3d_gemm_ws_bytes(int N, int K);
int t3d_gemm_f32_ws(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C,
                    int ldc, int M, int N, int K, int splitk, const float* bias, void* ws, size_t ws_bytes, t3d_stream_t stream);
int t3d_linear_f32_ws(const float* X, int ldx, const float* W, int ldw, const float* bias, const float* gbias,
                      int rows_per_group, float* Y, int ldy, int M, int K, int N, int act, const float* rowmask,
                      float* gmax, void* ws, size_t ws_bytes, t3d_stream_t stream);
int t3d_get_f32_engine(void);
/* mode 0: o0 = sum_r d, o1 = sum_r d^2 with d = x - y[c] when y != NULL (a per-column shift, e.g. row 0 of X: keeps the
 * variance E[d^2] - E[d]^2 free of cancellation), else d = x; mode 1: o0 = sum_r dy, o1 = sum_r dy*xhat with dy = X*act'(out), xhat=(y-mean)*rstd
 * (act = the T3D_ACT_* the forward applied after the batch norm) */
int t3d_colstats(const float* X, const float* out, const float* y, const float* mean, const float* rstd, float* o0,
                 float* o1, int M, int C, int mode, int act, t3d_stream_t stream);
/* sum / sumsq as produced by t3d_colstats mode 0 with the same `shift` (NULL: none) */
int t3d_bn_finalize(const float* sum, const float* sumsq, const float* shift, int M, int C, float eps, float decay, float* mean,
                    float* rstd, float* moving_mean, float* moving_var, t3d_stream_t stream);
int t3d_bn_apply(const float* y, const float* mean, const float* rstd, const float* gamma, const float* beta, float* out,
                 int M, int C, int act, t3d_stream_t stream);
int t3d_bn_backward(float* dOut, const float* out, const float* y, const float* mean, const float* rstd, const float* gamma,
                    const float* s1, const float* s2, int M, int C, int act, t3d_stream_t stream);
int t3d_maxpool_fwd(const float* x, int B, int N, int C, float* out, int* arg, t3d_stream_t stream);
int t3d_maxpool_bwd(const float* dout, const int* arg, int B, int N, int C, float* dx, t3d_stream_t stream);
/* the same with the `net * mask` of the masked stacks (semisup_models.py:184-185, 240-241) folded in: pools x * rowmask[row]
 * without materialising the product; the backward scales the routed gradient by the mask of the arg-max row */
int t3d_maxpool_masked_fwd(const float* x, const float* rowmask, int B, int N, int C, float* out, int* arg, t3d_stream_t stream);
int t3d_maxpool_masked_bwd(const float* dout, const int* arg, const float* rowmask, int B, int N, int C, float* dx,
                           t3d_stream_t stream);
/* model-A training step (train_semisup.py:199-262): soft_mask = softmax(logits)[1] (semisup_v1_sunrgbd.py:103); gradient of
 * sum_b w[b] * mean_n CE(logits[b,n,:], labels[b,n]) + (surface loss through soft_mask, gmask = d total / d soft_mask or
 * NULL) w.r.t. the mask logits; out[b,c] = sum_n x[b,n,c] (the gradient of conv6's per-frustum global half). */
int t3d_soft_mask(const float* logits, int B, int N, float* out, t3d_stream_t stream);
int t3d_seg_ce_bwd(const float* logits, const int* labels, const float* w, const float* gmask, int B, int N, float* dlogits,
                   t3d_stream_t stream);
int t3d_group_colsum(const float* x, int B, int N, int C, float* out, t3d_stream_t stream);
/* Row-sparse backward through a frozen (eval-mode) stack behind a max-pool (the BoxPC branch of train_semisup_adv.py:364-388):
 * the input gradient of the pooled layer is non-zero only in the rows that are the arg-max of some channel.  t3d_pool_rows:
 * per frustum the ascending list of distinct arg-max rows, padded with -1 to S >= min(C, N) slots (rows [B,S]), the slot of
 * every channel's arg-max row (slot [B,C]) and the number of distinct rows (count [B]); t3d_gather_rows: dst[b*S+s,:] =
 * src[b*N+rows[b,s],:] (zero rows for -1); t3d_scatter_pool_grad: dst [B*S,C] = 0, dst[b*S+slot[b,c], c] = g[b,c]. */
int t3d_pool_rows(const int* arg, int B, int N, int C, int S, int* rows, int* slot, int* count, t3d_stream_t stream);
int t3d_gather_rows(const float* src, const int* rows, int B, int N, int S, int C, float* dst, t3d_stream_t stream);
int t3d_scatter_pool_grad(const float* g, const int* slot, int B, int C, int S, float* dst, t3d_stream_t stream);
/* NORMALIZE_PC options of the BoxPC models (semisup_models.py:335-343, 413-421): pc [B,N,C] -> out [B,N,C] with xyz normalised
 * per cloud, channels >= 3 copied.  mode 0 'SD' = tf_normalize_point_clouds_to_mean_zero_and_unit_var (models/tf_util.py:157-173),
 * mode 1 'Spread' = tf_normalize_point_clouds_to_01 (:134-155). */
int t3d_normalize_pc(const float* pc, int B, int N, int C, int mode, float* out, t3d_stream_t stream);
int t3d_scale_mask(const float* x, const float* mask, float scale, float* out, long long n, t3d_stream_t stream);
/* ---- lazy batch norm (training-mode layers with M = B*N rows; tf_util.conv2d :1258-1323 + batch_norm_template :1645-1664) ----
 * The post-BN activation relu(gamma (y - mean) rstd + beta) of a layer is never written: consumers read the pre-BN tensor y
 * and apply the folded map relu(a_scale[c] * y + a_shift[c]) (a_scale = gamma rstd, a_shift = beta - mean a_scale, produced by
 * t3d_bn_finalize_affine) on the fly, and the batch statistics of y are accumulated by the GEMM that produces it.
 * t3d_gemm_bn_f32 = t3d_gemm_f32_ws with (a) a_scale / a_shift != NULL: A := relu(a_scale[c] * A + a_shift[c]), c = k for a
 * k-contiguous A (forward: needs the workspace path, K % 32 == 0, 16-byte aligned arrays), c = m for a row-contiguous A
 * (wgrad); (b) st_sum / st_sq != NULL: column sums of (C - st_shift[n]) and of its square (splitk == 1), zeroed inside.
 * Returns T3D_ERR_SHAPE when the problem does not reach a kernel that implements the extras; t3d_gemm_bn_supported says so
 * beforehand (kind 0 forward: 1 = tensor-core path, 2 = first-layer kernel (statistics only), 0 = no; kind 1 wgrad). */
int t3d_gemm_bn_supported(int M, int N, int K, int kind);
int t3d_gemm_bn_f32(const float* A, long long sam, long long sak, const float* a_scale, const float* a_shift, const float* B,
                    long long sbk, long long sbn, float* C, int ldc, int M, int N, int K, int splitk, const float* bias,
                    float* st_sum, float* st_sq, const float* st_shift, void* ws, size_t ws_bytes, t3d_stream_t stream);
/* Forward-only lazy BN layer whose output only feeds the max-pool over the pool_rows rows of each group (inst_seg conv5 in
 * the semi-supervised step, semisup_models.py:93-97 with the net frozen): the GEMM epilogue produces the batch statistics
 * and, per group, the column max and min of the pre-BN output as order-preserving keys; the [M, N] output is never written.
 * relu(a x + b) is monotone in x, so t3d_pool_bn_finish gives pooled = relu(a max + b) (a >= 0) or relu(a min + b).
 * Needs the tensor-core workspace path (M >= 4096, 32 <= K <= 2048, K % 32 == 0), pool_rows % 128 == 0, M % pool_rows == 0;
 * key buffers: [M / pool_rows, N] 32-bit words each. */
int t3d_gemm_bn_pool_f32(const float* A, long long lda, const float* a_scale, const float* a_shift, const float* W, int ldw, int M,
                         int N, int K, const float* bias, float* st_sum, float* st_sq, const float* st_shift, int pool_rows,
                         void* pool_max_keys, void* pool_min_keys, void* ws, size_t ws_bytes, t3d_stream_t stream);
int t3d_pool_bn_finish(const void* pool_max_keys, const void* pool_min_keys, const float* a_scale, const float* a_shift, int groups,
                       int C, float* out, t3d_stream_t stream);
/* y0[n] = sum_k a(k) W[k, n] + bias[n] for one row a (lazy BN applied when a_scale != NULL): the shift of the statistics */
int t3d_row0(const float* a, const float* a_scale, const float* a_shift, const float* W, int ldw, const float* bias, int K, int N,
             float* y0, t3d_stream_t stream);
int t3d_bn_finalize_affine(const float* sum, const float* sumsq, const float* shift, int M, int C, float eps, float decay,
                           const float* gamma, const float* beta, float* mean, float* rstd, float* a_scale, float* a_shift,
                           float* moving_mean, float* moving_var, t3d_stream_t stream);
/* t3d_colstats mode 1 / t3d_bn_backward of a lazy layer: the ReLU mask is a_scale * y + a_shift > 0 */
int t3d_colstats_lazy(const float* dOut, const float* y, const float* mean, const float* rstd, const float* a_scale,
                      const float* a_shift, float* s1, float* s2, int M, int C, t3d_stream_t stream);
int t3d_bn_backward_lazy(float* dOut, const float* y, const float* mean, const float* rstd, const float* gamma,
                         const float* a_scale, const float* a_shift, const float* s1, const float* s2, int M, int C,
                         t3d_stream_t stream);
int t3d_maxpool_lazy_fwd(const float* y, const float* a_scale, const float* a_shift, const float* rowmask, int B, int N, int C,
                         float* out, int* arg, t3d_stream_t stream);
/* t3d_maxpool_masked_fwd / t3d_maxpool_lazy_fwd with a caller-owned scratch of 8 * B * C bytes (8-byte aligned): the N rows of a
 * group are split across blocks and merged by a 64-bit atomic max (value, then smallest row) when B * C threads cannot fill
 * the GPU; keys == NULL is the serial kernel. */
int t3d_maxpool_fwd_ws(const float* x, const float* a_scale, const float* a_shift, const float* rowmask, int B, int N, int C,
                       float* out, int* arg, void* keys, t3d_stream_t stream);
/* backward of [BN -> ReLU -> (x rowmask) -> max-pool over the N rows of each group] from the pooled gradient g [B, C]: the BN
 * reductions are O(B C) gathers at the arg-max elements, dY [B*N, C] is one dense pass + a B x C scatter; s1 = d beta,
 * s2 = d gamma.  Replaces t3d_maxpool_masked_bwd + t3d_colstats + t3d_bn_backward (semisup_models.py:184-189, 240-245). */
int t3d_pool_bn_backward(const float* g, const int* arg, const float* rowmask, const float* y, const float* mean,
                         const float* rstd, const float* gamma, const float* a_scale, const float* a_shift, int B, int N, int C,
                         float* s1, float* s2, float* dY, t3d_stream_t stream);

typedef struct {
  const float *out9, *y_iou, *y_dc, *y_ds, *y_da;
  int B; float fit_bound, w_cls, w_delta, wc, ws, wa; int huber;
  float *cls_losses, *delta_losses, *total, *grad;
  /* class-confidence weighting (boxpc_sunrgbd.py:76-91, 166-177), p1 = softmax(fit logits)[1]:
   * pred_weigh 1: the deltas entering the loss are out9[0:7] * (1 - p1) (BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF; pass the RAW
   * network output); loss_weigh 1: delta loss * (1 - p1) (BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF), 2: * (1 - y_iou)
   * (..._BY_CLS_GT); stop_grad 1: p1 is a constant in both (BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA).  All 0: the plain loss. */
  int pred_weigh, loss_weigh, stop_grad;
} t3d_boxpc_loss_args;
int t3d_boxpc_loss(const t3d_boxpc_loss_args* args /* host */, t3d_stream_t stream);
/* theta -= lr_t * m / (sqrt(v) + eps), lr_t = lr*sqrt(1-b2^t)/(1-b1^t) supplied by the host (TF Adam) */
int t3d_adam(float* param, const float* grad, float* m, float* v, long long n, float lr_t, float beta1, float beta2,
             float eps, float grad_scale, t3d_stream_t stream);

/* ---- semi-supervised training step (train_semisup_adv.py:267-425): losses with gradients ---------------------
 * mean_N softmax cross-entropy of the mask logits per frustum (semisup_v1_sunrgbd.py:430-431) */
int t3d_seg_ce(const float* logits, const int* labels, int B, int N, float* out, t3d_stream_t stream);
/* per-class sums / counts of dims_reg over the batch (group means of weak_losses.get_intraclass_variance_loss_v1) */
int t3d_class_dims_stats(const float* dims_reg, const float* one_hot, int B, int NC, float* cls_sum, float* cls_cnt,
                         t3d_stream_t stream);
/* semisup_v1_sunrgbd.get_semi_loss_final (:323-421) = get_strong_loss(prefix F_) (:423-553) + weak_losses.get_reprojection_loss
 * (weak_losses.py:69-238) + get_intraclass_variance_loss_v1 (:267-291) + the BoxPC fit loss, with d total / d F_output,
 * d / d stage1_center, d / d F_pred_box_reg (g_reg, the part that flows through the regression-format box) and d / d fit_logits. */
typedef struct {
  const float *out, *stage1_center, *mask_losses, *one_hot;
  const float* y_center; const int* y_orient_cls; const float* y_orient_reg; const int* y_dims_cls; const float* y_dims_reg;
  const float *Rtilt, *K, *rot_frust, *box2D, *img_dim;
  const int* is_data_2D;
  const float *fit_logits, *mean_size, *cls_sum, *cls_cnt;
  const float* reg_in;   /* optional [B,7] (center, dims, orient): evaluate the weak losses on this box instead of the parsed one */
  int B, NH, NS, NC;
  unsigned icv_train_mask;
  float w_ce, box_mult, w_center, w_ocls, w_dcls, w_oreg, w_dreg, w_tnet, w_corner;
  float weak_mult, w_icv, w_reproj, w_fit;
  int reproj_only_2d, fit_only_2d, use_softmax_proj;
  float softmax_scale, dilate;
  int clip_lower_b, clip_pred_box, reproj_mse, icv_mse, train_box_mask;
  float inv_n3d;
  float *dF, *ds1, *g_reg, *dfit, *per_sample, *total;
} t3d_semi_loss_args;
int t3d_semi_loss(const t3d_semi_loss_args* args /* host */, t3d_stream_t stream);
/* dF += chain of g_reg [B,7] through tf_convert_box_params_from_anchor_to_reg_format (tf_util.py:1001-1041); ds1 += g_reg[:,0:3] */
int t3d_box_reg_backward(const float* out, const float* g_reg, const float* mean_size, int B, int NH, int NS, float* dF,
                         float* ds1, t3d_stream_t stream);
/* backward of tf_get_box_pc_representation (tf_util.py:764-795) w.r.t. the box: g6 [B*N,6] -> g_box [B,7] (+=) */
int t3d_boxpc_features_bwd(const float* pc, int B, int N, int C, const float* center, const float* orient, const float* g6,
                           float* g_box, t3d_stream_t stream);
int t3d_act_bwd(float* dout, const float* out, long long n, int act, t3d_stream_t stream);
int t3d_rowmask_mul(const float* x, const float* rowmask, float* out, long long M, int C, t3d_stream_t stream);
/* out[b,c] = scale * sum_n x[b,n,c], C <= 8 */
int t3d_group_sum(const float* x, int B, int N, int C, float scale, float* out, t3d_stream_t stream);

/* ---- inference post-processing ------------------------------------------------------
 * The numpy tail of test_semisup.inference (sunrgbd_detection/test_semisup.py:236-258) for one batch:
 *   pred_seg = argmax(logits, 2); mask_mean_prob = sum(softmax(logits)[:,:,1] * pred_seg) / (sum(pred_seg) + 1);
 *   score = log(mask_mean_prob + .01) + log(max softmax(heading_scores) + .01) + log(max softmax(size_scores) + .01)
 *           [+ log(fit_prob + .01)]; heading_cls / size_cls = argmax of the scores and the residuals they select.
 * pred_seg and mask_mean_prob may be NULL. */
typedef struct {
  const float *logits, *heading_scores, *heading_residuals, *size_scores, *size_residuals, *fit_prob;
  int B, N, NH, NS;
  unsigned char* pred_seg;
  float* mask_mean_prob;
  int* heading_cls;
  float* heading_res;
  int* size_cls;
  float* size_res;
  float* scores;
} t3d_infer_score_args;
int t3d_inference_scores(const t3d_infer_score_args* args /* host */, t3d_stream_t stream);
/* roi_seg_box3d_dataset.from_prediction_to_label_format (roi_seg_box3d_dataset.py:461-466, with class2angle :64-71,
 * class2size :79-82, rotate_pc_along_y :37-45) for a batch: out7[b] = (h, w, l, tx, ty, tz, ry). */
int t3d_prediction_to_label(const float* center, const int* heading_cls, const float* heading_res, const int* size_cls,
                            const float* size_res, const float* rot_angle, const float* mean_size /* [NS,3] (l,w,h) */,
                            int B, int NH, float* out7, t3d_stream_t stream);

/* ---- surface loss (first item) ----------------------------------------------------------
 * weak_losses.get_surface_loss (models/weak_losses.py:240-265) with tf_distance_to_closest_3D_box_surface_multi
 * (models/tf_util.py:610-720): loss[b] = mean_n max(0, d(p_bn, box_b) - margin) * soft_mask[b,n], d = the minimum of the six
 * (uncleaned, as the reference returns them) point-to-surface distances of the box (center, dims * scale_dims, orient).
 * With upstream != NULL (d total / d loss[b]) the same pass writes g_box [B,7] = d total / d (center, dims, orient) -- groups
 * gated by the train_* flags (WEAK_TRAIN_BOX_W_SURFACE) -- and g_mask [B,N] = d total / d soft_mask. */
typedef struct {
  const float* pc; int C;
  const float* soft_mask;
  const float *center, *dims, *orient;
  int B, N;
  float margin, scale_dims;
  int train_center, train_dims, train_orient;
  const float* upstream;
  float *loss, *g_box, *g_mask;
} t3d_surface_loss_args;
int t3d_surface_loss(const t3d_surface_loss_args* args /* host */, t3d_stream_t stream);

/* weak_losses.get_inactive_volume_loss_v1 (models/weak_losses.py:38-67) -> out[0]; with total / g_reg (the buffers of
 * t3d_semi_loss) the term is also folded in as get_semi_loss_final does (semisup_v1_sunrgbd.py:348-360):
 * total[4] += w iv, total[0] += mult w iv, g_reg[:,3:6] += mult w d iv / d dims.  Launch after t3d_semi_loss, same stream. */
int t3d_inactive_volume_loss(const float* dims, const float* one_hot, const float* margins, int B, int NC,
                             unsigned train_mask, float w, float mult, float* out, float* total, float* g_reg,
                             t3d_stream_t stream);

/* ---- frustum batch assembly (input side) ------------------------------------------------
 * ROISegBoxDataset.__getitem__ + get_batch (sunrgbd_detection/roi_seg_box3d_dataset.py:259-345, 370-417) for a batch:
 * resample every selected frustum to N points with the given choice indices, rotate to the centre view
 * (rot_angle = pi/2 + frustum_angle, rotate_pc_along_y), optional flip / shift augmentation, and -- when box3d != NULL --
 * the label encoding (box centre, angle2class, size2class).  The dataset is one flat device array of points + offsets. */
typedef struct {
  const float* points; int C_src;
  const int* labels;
  const long long* pt_off;
  const int *sel, *choice;
  const float *frustum_angle, *box3d, *heading, *size;
  const int* cls;
  const float* mean_size;
  const unsigned char* flip;
  const float *shift_z, *shift_y;
  int B, N, C_out, rotate_to_center, NH;
  float* batch_data;
  int* batch_label;
  float* center;
  int* heading_class; float* heading_residual;
  int* size_class; float* size_residual;
  float* rot_angle;
} t3d_assemble_args;
int t3d_assemble_frustum_batch(const t3d_assemble_args* args /* host */, t3d_stream_t stream);

/* Input wire format of the inference path: xyz fp32 [n,3] + rgb uint8 [n,3] (colours are k / 255 of 8-bit images in the
 * prepared SUN-RGBD frustums) -> the fp32 [n,6] layout of the reference's pc placeholder (semisup_v1_sunrgbd.py:39),
 * bit-identical to float32(k) / 255 on the host.  15 B instead of 24 B per point over PCIe. */
int t3d_assemble_points(const float* xyz, const uint8_t* rgb, long long n_points, float* out, t3d_stream_t stream);

/* ---- detection evaluation ------------------------------------------------------------
 * The matching loop of eval_det.eval_det_cls (sunrgbd_detection/eval_det.py:118-145) for one class: detections sorted by
 * descending score, 3D IoU (box_util.box3d_iou, the get_iou hook of eval_det.py:63-69) against the ground-truth boxes of
 * the same image, first-maximum / first-claim greedy matching -> tp, fp per detection (+ optional ovmax, jmax).
 * Detections and boxes are grouped by image through CSR offsets; gt_det is ng bytes of scratch. */
typedef struct {
  const float* det_corners;  /* [nd,8,3] */
  const int* img_det_off;    /* [nimg+1] */
  const int* img_det_idx;    /* [nd] positions in the sorted order, ascending inside an image */
  const float* gt_corners;   /* [ng,8,3] */
  const int* img_gt_off;     /* [nimg+1] */
  int nimg, nd, ng;
  float ovthresh;
  float *tp, *fp, *ovmax;
  int* jmax;
  unsigned char* gt_det;
} t3d_det_match_args;
int t3d_det_match(const t3d_det_match_args* args /* host */, t3d_stream_t stream);

/* Debug hook (not part of the reference-facing surface): install a device buffer of
 * 4 * 8192 uint64 into which CTA 0 of the tcgen05 kernels records (clock64 << 8 | tag) per role
 * (0 weight producer, 1 MMA issuer, 2 epilogue warp, 3 front warp); NULL switches tracing off. */
int t3d_set_trace_buffer(void* dev_buf);

#ifdef __cplusplus
}
#endif
#endif
