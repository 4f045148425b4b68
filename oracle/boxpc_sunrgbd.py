"""BoxPC-Fit model and losses: restatement of sunrgbd_detection/boxpc_sunrgbd.py:33-229."""
import torch

from . import semisup_models
from .weak_losses import _tf_huber, _tf_mse
from .semisup_v1_sunrgbd import convert_raw_y_box_to_reg_format  # noqa: F401 (identical copy, :206-229)
from transferable3d_b200.constants import NUM_CLASS


def placeholder_inputs(batch_size, num_point, num_channels):
    """boxpc_sunrgbd.py:33-54 -- NOTE the return order puts y_dims_delta before y_orient_delta."""
    f, i = 'float32', 'int32'
    B, N, C = batch_size, num_point, num_channels
    return (('pc', (B, N, C), f), ('one_hot_vec', (B, NUM_CLASS), f), ('y_seg', (B, N), i),
            ('x_center', (B, 3), f), ('x_orient_cls', (B,), i), ('x_orient_reg', (B,), f),
            ('x_dims_cls', (B,), i), ('x_dims_reg', (B, 3), f), ('y_box_iou', (B,), f),
            ('y_center_delta', (B, 3), f), ('y_dims_delta', (B, 3), f), ('y_orient_delta', (B,), f))


def get_model(boxpc, is_training, one_hot_vec, vs, use_one_hot_vec=False, bn_decay=None, c=None):
    """boxpc_sunrgbd.py:56-100."""
    end_points = {'class_ids': torch.argmax(one_hot_vec, dim=1).to(torch.int32)}
    box_reg, pc = boxpc
    delta_dims = 3 + 3 + 1
    if not use_one_hot_vec:
        one_hot_vec = None
    output, feats = semisup_models.box_pc_mask_features_model(
        box_reg, pc, None, 2 + delta_dims, is_training, end_points, False, False, vs,
        one_hot_vec=one_hot_vec, norm_box2D=None, bn_decay=bn_decay, c=c, scope='box_pc_mask_model')
    boxpc_fit_logits = output[:, -2:]
    logits_for_weigh = torch.softmax(boxpc_fit_logits, dim=1)[:, 1]
    pred_boxpc_fit = (torch.softmax(boxpc_fit_logits, dim=1)[:, 1] > 0.5).to(torch.int32)
    if c.BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA:
        logits_for_weigh = logits_for_weigh.detach()
    end_points['boxpc_feats_dict'] = feats
    end_points['boxpc_fit_logits'] = boxpc_fit_logits
    end_points['pred_boxpc_fit'] = pred_boxpc_fit
    end_points['logits_for_weigh'] = logits_for_weigh
    d_center, d_size, d_angle = output[:, 0:3], output[:, 3:6], output[:, 6]
    if c.BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF:
        w = 1. - logits_for_weigh
        d_center, d_size, d_angle = d_center * w.unsqueeze(1), d_size * w.unsqueeze(1), d_angle * w
    end_points['boxpc_delta_center'] = d_center
    end_points['boxpc_delta_size'] = d_size
    end_points['boxpc_delta_angle'] = d_angle
    return (boxpc_fit_logits, (d_center, d_size, d_angle)), end_points


def get_boxpc_cls_loss(logits, y_box_iou, end_points, reduce_loss=True, c=None):
    """boxpc_sunrgbd.py:130-141."""
    target = (y_box_iou > c.BOXPC_FIT_BOUNDS[0]).long()
    losses = torch.nn.functional.cross_entropy(logits, target, reduction='none')
    return losses.mean() if reduce_loss else losses


def get_boxpc_delta_loss(pred, labels, end_points, reduce_loss=True, c=None):
    """boxpc_sunrgbd.py:143-193."""
    logits, (d_center, d_size, d_angle) = pred
    y_box_iou, (y_center, y_size, y_angle) = labels
    lf = _tf_huber if c.BOXPC_DELTA_LOSS_TYPE == 'huber' else _tf_mse
    lc, ls, la = lf(y_center, d_center), lf(y_size, d_size), lf(y_angle, d_angle)
    assert not (c.BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF and c.BOXPC_WEIGH_DELTA_LOSS_BY_CLS_GT)
    w = 1.
    if c.BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF:
        w = 1. - end_points['logits_for_weigh']
    if c.BOXPC_WEIGH_DELTA_LOSS_BY_CLS_GT:
        w = 1. - y_box_iou
    lc, ls, la = lc.mean(dim=1) * w, ls.mean(dim=1) * w, la * w
    losses = (c.BOXPC_WEIGHT_DELTA_CENTER_PERCENT * lc + c.BOXPC_WEIGHT_DELTA_SIZE_PERCENT * ls +
              c.BOXPC_WEIGHT_DELTA_ANGLE_PERCENT * la)
    return losses.mean() if reduce_loss else losses


def get_loss(pred, labels, end_points, reduce_loss=True, c=None):
    """boxpc_sunrgbd.py:106-128."""
    logits, _ = pred
    y_box_iou, _ = labels
    cls_losses = get_boxpc_cls_loss(logits, y_box_iou, end_points, False, c)
    delta_losses = get_boxpc_delta_loss(pred, labels, end_points, False, c)
    total = c.BOXPC_WEIGHT_CLS * cls_losses + c.BOXPC_WEIGHT_DELTA * delta_losses
    return total.mean() if reduce_loss else total
