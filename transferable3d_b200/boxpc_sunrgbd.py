"""Mirror of sunrgbd/sunrgbd_detection/boxpc_sunrgbd.py on the B200 path: placeholder_inputs :33-54, get_model :56-100,
get_loss / get_boxpc_cls_loss / get_boxpc_delta_loss :106-193, convert_raw_y_box_to_reg_format :206-229."""
import ctypes

import numpy as np
import torch

from . import runtime as rt
from . import tf_util, semisup_models
from ._lib import ptr, stream, call, t3d_refine_args, t3d_boxpc_loss_args
from .constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER, NUM_CLASS, MEAN_DIMS_ARR, ORIENT_ANCHORS


def placeholder_inputs(batch_size, num_point, num_channels, device='cuda'):
    """boxpc_sunrgbd.py:33-54 -- NOTE y_dims_delta comes before y_orient_delta in the returned tuple."""
    f, i = torch.float32, torch.int32
    B, N, C = batch_size, num_point, num_channels
    Z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=device)
    return (Z((B, N, C), f), Z((B, NUM_CLASS), f), Z((B, N), i), Z((B, 3), f), Z((B,), i), Z((B,), f),
            Z((B,), i), Z((B, 3), f), Z((B,), f), Z((B, 3), f), Z((B, 3), f), Z((B,), f))


def parse_boxpc_output(output, c, curr_box=None, totals=None, weigh_during_test=False):
    """Output slicing of boxpc_sunrgbd.py:70-95 in one kernel; optionally applies one refine step of
    test_semisup.py:116-134 in place to curr_box / totals."""
    output = rt.f32(output)
    B = output.shape[0]
    dev = output.device
    E = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    o = dict(boxpc_fit_logits=E(B, 2), fit_prob=E(B), pred_boxpc_fit=torch.empty((B,), dtype=torch.int32, device=dev),
             boxpc_delta_center=E(B, 3), boxpc_delta_size=E(B, 3), boxpc_delta_angle=E(B))
    cb = curr_box if curr_box is not None else (None, None, None)
    tt = totals if totals is not None else (None, None, None)
    a = t3d_refine_args(ptr(output), B, int(bool(c.BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF)), int(bool(weigh_during_test)),
                        ptr(o['boxpc_fit_logits']), ptr(o['fit_prob']), ptr(o['pred_boxpc_fit']),
                        ptr(o['boxpc_delta_center']), ptr(o['boxpc_delta_size']), ptr(o['boxpc_delta_angle']),
                        ptr(cb[0]), ptr(cb[1]), ptr(cb[2]), ptr(tt[0]), ptr(tt[1]), ptr(tt[2]))
    call('t3d_boxpc_refine', ctypes.byref(a), stream())
    return o


def get_model(boxpc, is_training, one_hot_vec, use_one_hot_vec=False, bn_decay=None, c=None, _refine=None):
    """boxpc_sunrgbd.py:56-100 -> ((fit_logits, (dcenter, dsize, dangle)), end_points)."""
    end_points = {'class_ids': torch.argmax(one_hot_vec, dim=1).to(torch.int32)}
    box_reg, pc = boxpc
    delta_dims = 3 + 3 + 1
    if not use_one_hot_vec:
        one_hot_vec = None
    output, feats = semisup_models.box_pc_mask_features_model(box_reg, pc, None, 2 + delta_dims, is_training,
                                                              end_points=end_points, reuse=False, bn_for_output=False,
                                                              one_hot_vec=one_hot_vec, norm_box2D=None, bn_decay=bn_decay,
                                                              c=c, scope='box_pc_mask_model')
    kw = _refine or {}
    o = parse_boxpc_output(output, c, **kw)
    end_points['boxpc_feats_dict'] = feats
    end_points['boxpc_fit_logits'] = o['boxpc_fit_logits']
    end_points['pred_boxpc_fit'] = o['pred_boxpc_fit']
    end_points['logits_for_weigh'] = o['fit_prob']          # softmax(fit_logits)[:,1]
    end_points['boxpc_delta_center'] = o['boxpc_delta_center']
    end_points['boxpc_delta_size'] = o['boxpc_delta_size']
    end_points['boxpc_delta_angle'] = o['boxpc_delta_angle']
    pred_delta_box = (o['boxpc_delta_center'], o['boxpc_delta_size'], o['boxpc_delta_angle'])
    pred = (o['boxpc_fit_logits'], pred_delta_box)
    return pred, end_points


def convert_raw_y_box_to_reg_format(y_box, one_hot_vec):
    """boxpc_sunrgbd.py:206-229 (copy at semisup_v1_sunrgbd.py:584-608)."""
    y_centers, y_orient_cls, y_orient_reg, y_dims_cls, y_dims_reg = y_box
    st = rt.store()
    class_ids = torch.argmax(one_hot_vec, dim=1).to(torch.int32)
    dims_anchors = st.const('MEAN_DIMS_ARR', MEAN_DIMS_ARR)
    orient_anchors = st.const('ORIENT_ANCHORS', ORIENT_ANCHORS)
    Fn = torch.nn.functional
    dims_cls = Fn.one_hot(y_dims_cls.long(), NUM_SIZE_CLUSTER).to(torch.float32)
    dims_reg = tf_util.tf_expand_tile(rt.f32(y_dims_reg), axis=1, tile=[1, NUM_SIZE_CLUSTER, 1]).contiguous()
    orient_cls = Fn.one_hot(y_orient_cls.long(), NUM_HEADING_BIN).to(torch.float32)
    orient_reg = tf_util.tf_expand_tile(rt.f32(y_orient_reg), axis=1, tile=[1, NUM_HEADING_BIN]).contiguous()
    box = (y_centers, dims_cls, dims_reg, orient_cls, orient_reg)
    return tf_util.tf_convert_box_params_from_anchor_to_reg_format_multi(box, class_ids, dims_anchors, orient_anchors)


def _boxpc_losses(pred, labels, end_points, c):
    """One launch of the fused BoxPC loss kernel (csrc/train_ops.cuh, t3d_boxpc_loss): per-sample classification and delta
    losses, the scalar total of get_loss and d total / d (delta_center, delta_size, delta_angle, fit logits)."""
    logits, (d_center, d_size, d_angle) = pred
    y_box_iou, (y_center, y_size, y_angle) = labels
    assert not (c.BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF and c.BOXPC_WEIGH_DELTA_LOSS_BY_CLS_GT)
    # the deltas in `pred` are what get_model returned (already scaled when BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF); the loss
    # weight 1 - softmax(fit logits)[1] equals end_points['logits_for_weigh'] (boxpc_sunrgbd.py:166-177)
    loss_weigh = 1 if c.BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF else (2 if c.BOXPC_WEIGH_DELTA_LOSS_BY_CLS_GT else 0)
    B, dev = logits.shape[0], logits.device
    out9 = torch.cat([rt.f32(d_center).reshape(B, 3), rt.f32(d_size).reshape(B, 3), rt.f32(d_angle).reshape(B, 1),
                      rt.f32(logits).reshape(B, 2)], dim=1).contiguous()
    y_iou, y_dc, y_ds, y_da = rt.f32(y_box_iou), rt.f32(y_center), rt.f32(y_size), rt.f32(y_angle)
    E = lambda *s_: torch.empty(s_, dtype=torch.float32, device=dev)
    cls_l, del_l, total, g9 = E(B), E(B), E(1), E(B, 9)
    a = t3d_boxpc_loss_args(ptr(out9), ptr(y_iou), ptr(y_dc), ptr(y_ds), ptr(y_da), B, float(c.BOXPC_FIT_BOUNDS[0]),
                            float(c.BOXPC_WEIGHT_CLS), float(c.BOXPC_WEIGHT_DELTA), float(c.BOXPC_WEIGHT_DELTA_CENTER_PERCENT),
                            float(c.BOXPC_WEIGHT_DELTA_SIZE_PERCENT), float(c.BOXPC_WEIGHT_DELTA_ANGLE_PERCENT),
                            1 if c.BOXPC_DELTA_LOSS_TYPE == 'huber' else 0, ptr(cls_l), ptr(del_l), ptr(total), ptr(g9),
                            0, loss_weigh, 1 if c.BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA else 0)
    call('t3d_boxpc_loss', ctypes.byref(a), stream())
    if end_points is not None:
        end_points['boxpc_loss_grad'] = g9        # columns: delta_center 0:3, delta_size 3:6, delta_angle 6, fit logits 7:9
    return cls_l, del_l, total


def get_boxpc_cls_loss(logits, y_box_iou, end_points, reduce_loss=True, c=None):
    """boxpc_sunrgbd.py:130-141: softmax cross-entropy against (iou > BOXPC_FIT_BOUNDS[0]) -> (B,) (mean if reduce_loss)."""
    B, dev = logits.shape[0], logits.device
    z = lambda *s_: torch.zeros(s_, dtype=torch.float32, device=dev)
    cls_l, _, _ = _boxpc_losses((logits, (z(B, 3), z(B, 3), z(B))), (y_box_iou, (z(B, 3), z(B, 3), z(B))), None, c)
    return cls_l.mean() if reduce_loss else cls_l


def get_boxpc_delta_loss(pred, labels, end_points, reduce_loss=True, c=None):
    """boxpc_sunrgbd.py:143-193: weighted huber / mse of the three box deltas -> (B,) (mean if reduce_loss)."""
    _, del_l, _ = _boxpc_losses(pred, labels, None, c)
    return del_l.mean() if reduce_loss else del_l


def get_loss(pred, labels, end_points, reduce_loss=True, c=None):
    """boxpc_sunrgbd.py:106-128: BOXPC_WEIGHT_CLS * cls + BOXPC_WEIGHT_DELTA * delta -> scalar mean (or (B,))."""
    cls_l, del_l, total = _boxpc_losses(pred, labels, end_points, c)
    if reduce_loss:
        return total[0]
    return float(c.BOXPC_WEIGHT_CLS) * cls_l + float(c.BOXPC_WEIGHT_DELTA) * del_l
