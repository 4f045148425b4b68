"""ctypes binding of libt3d_b200.so (C ABI in include/t3d_b200.h).

There is no CPU or library fallback: if the shared object is missing, or a call is made with
tensors that are not on a CUDA device, this module raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libt3d_b200.so')

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int


class t3d_parse_args(_c.Structure):
    _fields_ = [('output', _P), ('stage1_center', _P), ('mean_size', _P), ('orient_anchors', _P),
                ('B', _I), ('NH', _I), ('NS', _I),
                ('center', _P), ('heading_scores', _P), ('heading_res_norm', _P), ('heading_res', _P),
                ('size_scores', _P), ('size_res_norm', _P), ('size_res', _P),
                ('reg_center', _P), ('reg_dims', _P), ('reg_orient', _P)]


class t3d_refine_args(_c.Structure):
    _fields_ = [('out9', _P), ('B', _I), ('weigh_pred_by_conf', _I), ('weigh_during_test', _I),
                ('fit_logits', _P), ('fit_prob', _P), ('pred_fit', _P),
                ('delta_center', _P), ('delta_size', _P), ('delta_angle', _P),
                ('box_center', _P), ('box_dims', _P), ('box_orient', _P),
                ('tot_center', _P), ('tot_size', _P), ('tot_angle', _P)]


class t3d_boxpc_loss_args(_c.Structure):
    _fields_ = [('out9', _P), ('y_iou', _P), ('y_dc', _P), ('y_ds', _P), ('y_da', _P),
                ('B', _I), ('fit_bound', _c.c_float), ('w_cls', _c.c_float), ('w_delta', _c.c_float),
                ('wc', _c.c_float), ('ws', _c.c_float), ('wa', _c.c_float), ('huber', _I),
                ('cls_losses', _P), ('delta_losses', _P), ('total', _P), ('grad', _P),
                ('pred_weigh', _I), ('loss_weigh', _I), ('stop_grad', _I)]


class t3d_semi_loss_args(_c.Structure):
    _fields_ = ([(n, _P) for n in ('out', 'stage1_center', 'mask_losses', 'one_hot', 'y_center', 'y_orient_cls', 'y_orient_reg',
                                   'y_dims_cls', 'y_dims_reg', 'Rtilt', 'K', 'rot_frust', 'box2D', 'img_dim', 'is_data_2D',
                                   'fit_logits', 'mean_size', 'cls_sum', 'cls_cnt', 'reg_in')] +
                [(n, _I) for n in ('B', 'NH', 'NS', 'NC')] + [('icv_train_mask', _c.c_uint)] +
                [(n, _c.c_float) for n in ('w_ce', 'box_mult', 'w_center', 'w_ocls', 'w_dcls', 'w_oreg', 'w_dreg', 'w_tnet',
                                           'w_corner', 'weak_mult', 'w_icv', 'w_reproj', 'w_fit')] +
                [(n, _I) for n in ('reproj_only_2d', 'fit_only_2d', 'use_softmax_proj')] +
                [('softmax_scale', _c.c_float), ('dilate', _c.c_float)] +
                [(n, _I) for n in ('clip_lower_b', 'clip_pred_box', 'reproj_mse', 'icv_mse', 'train_box_mask')] +
                [('inv_n3d', _c.c_float)] +
                [(n, _P) for n in ('dF', 'ds1', 'g_reg', 'dfit', 'per_sample', 'total')])


class t3d_compute_iou_args(_c.Structure):
    _fields_ = ([(n, _P) for n in ('center_pred', 'heading_logits', 'heading_residuals', 'size_logits', 'size_residuals', 'center_label',
                                   'heading_class_label', 'heading_residual_label', 'size_class_label', 'size_residual_label', 'mean_size')] +
                [('B', _I), ('NH', _I), ('NS', _I), ('iou2ds', _P), ('iou3ds', _P)])


class t3d_perturb_args(_c.Structure):
    _fields_ = ([(n, _P) for n in ('center', 'size', 'heading', 'bounds')] + [('B', _I), ('max_attempts', _I)] +
                [(n, _c.c_float) for n in ('center_perturbation', 'size_perturbation', 'angle_perturbation')] + [('seed', _c.c_uint64)] +
                [(n, _P) for n in ('new_center', 'new_size', 'new_heading', 'iou3d', 'd_center', 'd_size', 'd_angle', 'attempts')])


class t3d_infer_score_args(_c.Structure):
    _fields_ = ([(n, _P) for n in ('logits', 'heading_scores', 'heading_residuals', 'size_scores', 'size_residuals', 'fit_prob')] +
                [(n, _I) for n in ('B', 'N', 'NH', 'NS')] +
                [(n, _P) for n in ('pred_seg', 'mask_mean_prob', 'heading_cls', 'heading_res', 'size_cls', 'size_res', 'scores')])


class t3d_surface_loss_args(_c.Structure):
    _fields_ = [('pc', _P), ('C', _I), ('soft_mask', _P), ('center', _P), ('dims', _P), ('orient', _P), ('B', _I), ('N', _I),
                ('margin', _c.c_float), ('scale_dims', _c.c_float), ('train_center', _I), ('train_dims', _I), ('train_orient', _I),
                ('upstream', _P), ('loss', _P), ('g_box', _P), ('g_mask', _P)]


class t3d_assemble_args(_c.Structure):
    _fields_ = [('points', _P), ('C_src', _I), ('labels', _P), ('pt_off', _P), ('sel', _P), ('choice', _P), ('frustum_angle', _P),
                ('box3d', _P), ('heading', _P), ('size', _P), ('cls', _P), ('mean_size', _P), ('flip', _P), ('shift_z', _P),
                ('shift_y', _P), ('B', _I), ('N', _I), ('C_out', _I), ('rotate_to_center', _I), ('NH', _I), ('batch_data', _P),
                ('batch_label', _P), ('center', _P), ('heading_class', _P), ('heading_residual', _P), ('size_class', _P),
                ('size_residual', _P), ('rot_angle', _P)]


class t3d_det_match_args(_c.Structure):
    _fields_ = ([(n, _P) for n in ('det_corners', 'img_det_off', 'img_det_idx', 'gt_corners', 'img_gt_off')] +
                [(n, _I) for n in ('nimg', 'nd', 'ng')] + [('ovthresh', _c.c_float)] +
                [(n, _P) for n in ('tp', 'fp', 'ovmax', 'jmax', 'gt_det')])


_L = _c.c_longlong
_F = _c.c_float

# name -> (restype, argtypes); every symbol declared in include/t3d_b200.h
SIGNATURES = {
    't3d_version': (_I, []),
    't3d_error_string': (_c.c_char_p, [_I]),
    't3d_linear_f32': (_I, [_P, _I, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    't3d_mask_centroid': (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    't3d_resample': (_I, [_P, _P, _I, _I, _I, _I, _c.c_uint64, _P, _P, _P, _I, _P, _I, _P, _P]),
    't3d_build_tiles': (_I, [_P, _I, _I, _P, _P, _P]),
    't3d_prepare_xyz': (_I, [_P, _I, _I, _I, _P, _P, _P]),
    't3d_boxpc_features': (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P]),
    't3d_parse_box': (_I, [_c.POINTER(t3d_parse_args), _P]),
    't3d_anchor_to_reg': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    't3d_boxpc_refine': (_I, [_c.POINTER(t3d_refine_args), _P]),
    't3d_f2': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    't3d_box3d_corners_helper': (_I, [_P, _P, _P, _I, _P, _P]),
    't3d_box3d_corners_all': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P]),
    't3d_chain_arena_bytes': (_c.c_size_t, [_I]),
    't3d_chain_num_layers': (_I, [_I]),
    't3d_chain_tile_points': (_I, [_I]),
    't3d_chain_out_channels': (_I, [_I]),
    't3d_pack_chain': (_I, [_I, _c.POINTER(_P), _c.POINTER(_P), _P, _P]),
    't3d_chain_max_bf16': (_I, [_I, _P, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    't3d_chain_max_bf16_wire': (_I, [_I, _P, _P, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    't3d_seg2_arena_bytes': (_c.c_size_t, []),
    't3d_pack_seg2': (_I, [_P] * 10 + [_P]),
    't3d_seg_stage2_bf16': (_I, [_P, _P, _P, _P, _I, _I, _P]),
    't3d_chain_arena_bytes_x2': (_c.c_size_t, [_I]),
    't3d_pack_chain_x2': (_I, [_I, _c.POINTER(_P), _c.POINTER(_P), _c.POINTER(_F), _P, _P]),
    't3d_chain_max_x2': (_I, [_I, _P, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    't3d_seg2_arena_bytes_x2': (_c.c_size_t, []),
    't3d_pack_seg2_x2': (_I, [_P] * 9 + [_c.POINTER(_F), _P, _P]),
    't3d_seg_stage2_x2': (_I, [_P, _P, _P, _P, _I, _I, _P]),
    't3d_set_x2_debias': (_I, [_F]),
    't3d_get_x2_debias': (_F, []),
    't3d_set_trace_buffer': (_I, [_P]),
    't3d_gemm_f32': (_I, [_P, _L, _L, _P, _L, _L, _P, _I, _I, _I, _I, _I, _P, _P]),
    't3d_gemm_ws_bytes': (_c.c_size_t, [_I, _I]),
    't3d_gemm_f32_ws': (_I, [_P, _L, _L, _P, _L, _L, _P, _I, _I, _I, _I, _I, _P, _P, _c.c_size_t, _P]),
    't3d_linear_f32_ws': (_I, [_P, _I, _P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _c.c_size_t, _P]),
    't3d_set_f32_engine': (_I, [_I]),
    't3d_get_f32_engine': (_I, []),
    't3d_colstats': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    't3d_bn_finalize': (_I, [_P, _P, _P, _I, _I, _F, _F, _P, _P, _P, _P, _P]),
    't3d_bn_apply': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    't3d_bn_backward': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    't3d_gemm_bn_supported': (_I, [_I, _I, _I, _I]),
    't3d_gemm_bn_f32': (_I, [_P, _L, _L, _P, _P, _P, _L, _L, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _c.c_size_t, _P]),
    't3d_gemm_bn_pool_f32': (_I, [_P, _L, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P, _P, _c.c_size_t, _P]),
    't3d_pool_bn_finish': (_I, [_P, _P, _P, _P, _I, _I, _P, _P]),
    't3d_row0': (_I, [_P, _P, _P, _P, _I, _P, _I, _I, _P, _P]),
    't3d_bn_finalize_affine': (_I, [_P, _P, _P, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    't3d_colstats_lazy': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    't3d_bn_backward_lazy': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    't3d_maxpool_lazy_fwd': (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    't3d_maxpool_fwd_ws': (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    't3d_soft_mask': (_I, [_P, _I, _I, _P, _P]),
    't3d_seg_ce_bwd': (_I, [_P, _P, _P, _P, _I, _I, _P, _P]),
    't3d_group_colsum': (_I, [_P, _I, _I, _I, _P, _P]),
    't3d_pool_rows': (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P]),
    't3d_gather_rows': (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    't3d_scatter_pool_grad': (_I, [_P, _P, _I, _I, _I, _P, _P]),
    't3d_normalize_pc': (_I, [_P, _I, _I, _I, _I, _P, _P]),
    't3d_pool_bn_backward': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    't3d_maxpool_fwd': (_I, [_P, _I, _I, _I, _P, _P, _P]),
    't3d_maxpool_bwd': (_I, [_P, _P, _I, _I, _I, _P, _P]),
    't3d_maxpool_masked_fwd': (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    't3d_maxpool_masked_bwd': (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    't3d_scale_mask': (_I, [_P, _P, _F, _P, _L, _P]),
    't3d_boxpc_loss': (_I, [_c.POINTER(t3d_boxpc_loss_args), _P]),
    't3d_adam': (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _P]),
    't3d_seg_ce': (_I, [_P, _P, _I, _I, _P, _P]),
    't3d_class_dims_stats': (_I, [_P, _P, _I, _I, _P, _P, _P]),
    't3d_semi_loss': (_I, [_c.POINTER(t3d_semi_loss_args), _P]),
    't3d_box_reg_backward': (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    't3d_boxpc_features_bwd': (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P]),
    't3d_act_bwd': (_I, [_P, _P, _L, _I, _P]),
    't3d_rowmask_mul': (_I, [_P, _P, _P, _L, _I, _P]),
    't3d_group_sum': (_I, [_P, _I, _I, _I, _F, _P, _P]),
    't3d_get_3d_box': (_I, [_P, _P, _P, _I, _P, _P]),
    't3d_box3d_iou': (_I, [_P, _P, _I, _P, _P, _P]),
    't3d_compute_box3d_iou': (_I, [_c.POINTER(t3d_compute_iou_args), _P]),
    't3d_perturb_boxes': (_I, [_c.POINTER(t3d_perturb_args), _P]),
    't3d_inactive_volume_loss': (_I, [_P, _P, _P, _I, _I, _c.c_uint, _F, _F, _P, _P, _P, _P]),
    't3d_surface_loss': (_I, [_c.POINTER(t3d_surface_loss_args), _P]),
    't3d_assemble_frustum_batch': (_I, [_c.POINTER(t3d_assemble_args), _P]),
    't3d_det_match': (_I, [_c.POINTER(t3d_det_match_args), _P]),
    't3d_assemble_points': (_I, [_P, _P, _L, _P, _P]),
    't3d_inference_scores': (_I, [_c.POINTER(t3d_infer_score_args), _P]),
    't3d_prediction_to_label': (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P]),
}

_lib = None


class T3DError(RuntimeError):
    pass


def load():
    """Loads the shared library (no compute). Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise T3DError('%s is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                           '(nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        # 'simt' = CUDA-core SGEMM, 'tc' = tcgen05 bf16 x 3 (default), 'tc2' = bf16 x 2 (three products), 'bf16' = one pass
        engine = os.environ.get('T3D_F32_ENGINE')
        if engine is not None:
            codes = {'simt': 0, 'tc': 1, 'bf16': 2, 'tc2': 3}
            if engine not in codes:
                raise T3DError("T3D_F32_ENGINE must be one of %s" % sorted(codes))
            lib.t3d_set_f32_engine(codes[engine])
        _lib = lib
    return _lib


def check(code):
    if code != 0:
        msg = load().t3d_error_string(int(code))
        raise (ValueError if code < 0 else T3DError)('t3d call failed (%d): %s' % (code, msg.decode()))


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise T3DError('t3d_b200 ops need CUDA tensors (no CPU fallback)')
    if not t.is_contiguous():
        raise ValueError('t3d_b200 ops need contiguous tensors')
    return _P(t.data_ptr())


def stream():
    return _P(torch.cuda.current_stream().cuda_stream)


_WS_CALLS = ('t3d_gemm_f32', 't3d_linear_f32')
_WS_BYTES = 16 << 20          # >= t3d_gemm_ws_bytes(1152, 2048): every layer of the reference's networks
_workspaces = {}


def gemm_workspace():
    """Per (device, stream) scratch for the pre-split operand of the tensor-core GEMMs (caller-owned: the library never
    allocates)."""
    key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        ws = _workspaces[key] = torch.empty(_WS_BYTES, dtype=torch.uint8, device='cuda')
    return ws


def call(name, *args):
    if name in _WS_CALLS:
        ws = gemm_workspace()
        args = args[:-1] + (ptr(ws), ws.numel(), args[-1])
        name += '_ws'
    check(getattr(load(), name)(*args))
