"""Mirror of the hot-path part of models/tf_util.py (same function names / argument meaning),
running on the B200 through libt3d_b200.so.  Tensors are torch CUDA tensors; per-point maps
are (B,N,C) (the reference's NHWC singleton axis (B,N,1,C) is dropped).

Layer wrappers (tf_util.py:1258-1323, 1463-1524, 1720-1741) evaluate one layer in fp32 with BN
folded; the fused tcgen05 chains live in semisup_models / model_util where whole stacks are known.
"""
import numpy as np
import torch

from . import runtime as rt
from ._lib import ptr, stream, call, t3d_parse_args
import ctypes


def _act_name(activation_fn):
    if activation_fn is None:
        return None
    if activation_fn in ('relu', 'leaky_relu', 'tanh'):
        return activation_fn
    name = getattr(activation_fn, '__name__', str(activation_fn))
    if name in ('relu', 'leaky_relu', 'tanh'):
        return name
    raise ValueError('unsupported activation_fn %r' % (activation_fn,))


def conv2d(inputs, num_output_channels, kernel_size, scope, stride=[1, 1], padding='SAME', data_format='NHWC', use_xavier=True,
           stddev=1e-3, weight_decay=None, activation_fn='relu', bn=False, bn_decay=None, is_training=None):
    """tf_util.conv2d (tf_util.py:1258-1323; same positional order) for the 1x1 / [1,D] kernels of the hot path.
    inputs (B,N,Cin) -> (B,N,Cout).  use_xavier / stddev / weight_decay describe how the reference CREATES the variable; here the
    weights come from the variable store, so they are accepted and unused."""
    if data_format != 'NHWC':
        raise ValueError('conv2d: only NHWC (points x channels) is on the hot path')
    rt.require_eval(is_training)
    st = rt.store()
    layer = st.scope_name(scope)
    w, b = st.folded(layer)
    B, N, K = inputs.shape
    if w.shape != (K, num_output_channels):
        raise ValueError('%s: weights %s do not match input %s' % (layer, tuple(w.shape), tuple(inputs.shape)))
    y, _ = rt.linear(inputs.reshape(B * N, K), w, b, _act_name(activation_fn))
    return y.reshape(B, N, num_output_channels)


def fully_connected(inputs, num_outputs, scope, use_xavier=True, stddev=1e-3, weight_decay=None, activation_fn='relu', bn=False,
                    bn_decay=None, is_training=None):
    """tf_util.fully_connected (tf_util.py:1463-1499; same positional order). inputs (B,Cin).  The initialiser arguments are
    accepted and unused (weights come from the variable store)."""
    rt.require_eval(is_training)
    st = rt.store()
    layer = st.scope_name(scope)
    w, b = st.folded(layer)
    if w.shape != (inputs.shape[1], num_outputs):
        raise ValueError('%s: weights %s do not match input %s' % (layer, tuple(w.shape), tuple(inputs.shape)))
    y, _ = rt.linear(inputs, w, b, _act_name(activation_fn))
    return y


def max_pool2d(inputs, kernel_size, scope=None, stride=[2, 2], padding='VALID'):
    """tf_util.max_pool2d (tf_util.py:1501-1524) as the hot path uses it: kernel [num_point, 1] over the point axis.
    inputs (B,N,C) -> (B,1,C).  (Inside the fused chains the max is taken in the last layer's epilogue instead.)"""
    B, N, C = inputs.shape
    if int(kernel_size[0]) != N or int(kernel_size[1]) != 1:
        raise ValueError('max_pool2d: only the [num_point, 1] pooling of the PointNet stacks is on the hot path')
    x = rt.f32(inputs).reshape(B * N, C)
    pooled = torch.empty((B, C), dtype=torch.float32, device=x.device)
    arg = torch.empty((B, C), dtype=torch.int32, device=x.device)
    call('t3d_maxpool_fwd', ptr(x), B, N, C, ptr(pooled), ptr(arg), stream())
    return pooled.reshape(B, 1, C)


def dropout(inputs, is_training, scope, keep_prob=0.5, noise_shape=None):
    """tf_util.dropout (tf_util.py:1720-1741): identity when not training."""
    rt.require_eval(is_training)
    return inputs


def tf_expand_tile(tensor, axis, tile):
    """tf_util.py:1141-1142."""
    return tensor.unsqueeze(axis).repeat(*tile)


def tf_normalize_2D_bboxes(box2D, image_dim):
    """tf_util.py:466-484 (O(B) plumbing; only consumed when USE_NORMALIZED_BOX2D_AS_FEATS)."""
    rows, cols = image_dim[:, 0], image_dim[:, 1]
    return torch.stack([box2D[:, 0] / cols, box2D[:, 1] / rows, box2D[:, 2] / cols, box2D[:, 3] / rows], dim=1)


def _normalize_pc(point_clouds, mode):
    pc = rt.f32(point_clouds).contiguous()
    B, N, C = pc.shape
    out = torch.empty_like(pc)
    call('t3d_normalize_pc', ptr(pc), B, N, C, mode, ptr(out), stream())
    return out


def tf_normalize_point_clouds_to_mean_zero_and_unit_var(point_clouds):
    """tf_util.py:157-173: xyz -> (x - mean) / (sqrt(var) + 1e-5) per cloud; channels >= 3 kept."""
    return _normalize_pc(point_clouds, 0)


def tf_normalize_point_clouds_to_01(point_clouds):
    """tf_util.py:134-155: xyz -> (x - mean) / (largest xyz extent + 1e-5) per cloud; channels >= 3 kept."""
    return _normalize_pc(point_clouds, 1)


def tf_get_box_pc_representation(box_reg, pc):
    """tf_util.py:764-795: (B,N,C) -> (B,N,C+6) = pc ++ 6 signed plane distances."""
    center, dims, orient = [rt.f32(t) for t in box_reg]
    pc = rt.f32(pc)
    B, N, C = pc.shape
    out = torch.empty((B, N, C + 6), dtype=torch.float32, device=pc.device)
    call('t3d_boxpc_features', ptr(pc), B, N, C, ptr(center), ptr(dims), ptr(orient), ptr(out), stream())
    return out


def tf_convert_box_params_from_anchor_to_reg_format_multi(box_params, y_classes, dims_anchors, orient_anchors):
    """tf_util.py:1001-1041: argmax-select the anchors/residuals -> (center (B,3), dims (B,3), orient (B,))."""
    center, dims_cls, dims_reg, orient_cls, orient_reg = [rt.f32(t) for t in box_params]
    B, NS = dims_cls.shape
    NH = orient_cls.shape[1]
    dev = center.device
    oc = torch.empty((B, 3), dtype=torch.float32, device=dev)
    od = torch.empty((B, 3), dtype=torch.float32, device=dev)
    oo = torch.empty((B,), dtype=torch.float32, device=dev)
    da, oa = rt.f32(dims_anchors), rt.f32(orient_anchors)       # keep alive until the launch
    call('t3d_anchor_to_reg', ptr(center), ptr(dims_cls), ptr(dims_reg), ptr(orient_cls), ptr(orient_reg),
         ptr(da), ptr(oa), B, NS, NH, ptr(oc), ptr(od), ptr(oo), stream())
    return oc, od, oo


def parse_box_output(output, stage1_center, mean_size, orient_anchors, want_reg=True):
    """Fused slicing of the (B, 3+2NH+4NS) head output + anchor->reg (one kernel).
    Returns dict with center, heading_scores, heading_residuals_normalized, heading_residuals,
    size_scores, size_residuals_normalized, size_residuals [, reg = (center, dims, orient)]."""
    output = rt.f32(output)
    B = output.shape[0]
    NH, NS = orient_anchors.shape[0], mean_size.shape[0]
    assert output.shape[1] == 3 + 2 * NH + 4 * NS
    dev = output.device
    E = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    o = dict(center=E(B, 3), heading_scores=E(B, NH), heading_residuals_normalized=E(B, NH),
             heading_residuals=E(B, NH), size_scores=E(B, NS), size_residuals_normalized=E(B, NS, 3),
             size_residuals=E(B, NS, 3))
    reg = (E(B, 3), E(B, 3), E(B)) if want_reg else (None, None, None)
    a = t3d_parse_args(ptr(output), ptr(stage1_center), ptr(mean_size), ptr(orient_anchors), B, NH, NS,
                       ptr(o['center']), ptr(o['heading_scores']), ptr(o['heading_residuals_normalized']),
                       ptr(o['heading_residuals']), ptr(o['size_scores']), ptr(o['size_residuals_normalized']),
                       ptr(o['size_residuals']), ptr(reg[0]), ptr(reg[1]), ptr(reg[2]))
    call('t3d_parse_box', ctypes.byref(a), stream())
    if want_reg:
        o['reg'] = reg
    return o
