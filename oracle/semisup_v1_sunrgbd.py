"""Model + loss assembly: restatement of sunrgbd_detection/semisup_v1_sunrgbd.py:37-608.

get_iou_summary (:236-246) is a metrics-only tf.py_func around the missing box_util.box3d_iou
and is not restated (SURVEY 8f 'next').  get_semi_loss_backbone's surface loss is 'next' too.
"""
import numpy as np
import torch

from . import tf_util, weak_losses, semisup_models
from .model_util import get_box3d_corners_sunrgbd, get_box3d_corners_helper
from .tf_layers import leaky_relu
from transferable3d_b200.constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER, NUM_CLASS, MEAN_DIMS_ARR


def placeholder_inputs(batch_size, num_point, num_channel):
    """semisup_v1_sunrgbd.py:37-67: (name, shape, dtype) of the 18 placeholders, in order."""
    f, i = 'float32', 'int32'
    B, N, C = batch_size, num_point, num_channel
    return (('pc', (B, N, C), f), ('bg_pc', (B, N, C), f), ('img', (B, None, None, 3), f),
            ('one_hot_vec', (B, NUM_CLASS), f), ('labels', (B, N), i), ('centers', (B, 3), f),
            ('y_orient_cls', (B,), i), ('y_orient_reg', (B,), f), ('y_dims_cls', (B,), i),
            ('y_dims_reg', (B, 3), f), ('R0_rect', (B, 3, 3), f), ('P', (B, 3, 4), f),
            ('Rtilt', (B, 3, 3), f), ('K', (B, 3, 3), f), ('rot_frust', (B, 1), f),
            ('box2D', (B, 4), f), ('img_dim', (B, 2), f), ('is_data_2D', (B,), i))


def _base_end_points(pc, one_hot_vec):
    dt = pc.dtype
    return {'point_cloud': pc, 'class_one_hot': one_hot_vec,
            'class_ids': torch.argmax(one_hot_vec, dim=1).to(torch.int32),
            'dims_anchors': torch.as_tensor(MEAN_DIMS_ARR, dtype=torch.float32).to(dt),
            'orient_anchors': torch.as_tensor(np.arange(0, 2 * np.pi, 2 * np.pi / NUM_HEADING_BIN),
                                              dtype=torch.float32).to(dt)}


def get_semi_model(pc, bg_pc, img, one_hot_vec, is_training, use_one_hot, vs, oracle_mask=None,
                   norm_box2D=None, bn_decay=None, c=None):
    """semisup_v1_sunrgbd.py:69-79."""
    if c.SEMI_MODEL == 'A':
        return get_semi_model_backbone(pc, bg_pc, img, one_hot_vec, is_training, use_one_hot, vs,
                                       oracle_mask, norm_box2D, bn_decay, c)
    elif c.SEMI_MODEL == 'F':
        return get_semi_model_final(pc, bg_pc, img, one_hot_vec, is_training, use_one_hot, vs,
                                    oracle_mask, norm_box2D, bn_decay, c)
    raise Exception('Not implemented SEMI_MODEL: %s' % c.SEMI_MODEL)


def get_semi_model_backbone(pc, bg_pc, img, one_hot_vec, is_training, use_one_hot, vs, oracle_mask=None,
                            norm_box2D=None, bn_decay=None, c=None):
    """semisup_v1_sunrgbd.py:81-130 (model A)."""
    end_points = _base_end_points(pc, one_hot_vec)
    if not use_one_hot:
        one_hot_vec = None
    if oracle_mask is not None:
        raise NotImplementedError
    if not c.USE_NORMALIZED_BOX2D_AS_FEATS:
        norm_box2D = None
    logits = semisup_models.v1_inst_seg(pc, None, one_hot_vec, end_points, is_training, vs, bn_decay, 'inst_seg')
    end_points['soft_mask'] = torch.softmax(logits, dim=-1)[:, :, 1]
    mask, mean, pc_xyz, pc_xyz_stage1 = semisup_models.subtract_points_mean(pc, logits)
    stage1_center = semisup_models.v1_tnet(pc_xyz_stage1, mask, mean, one_hot_vec, end_points, is_training, vs,
                                           norm_box2D, bn_decay, 'tnet')
    pc_xyz_submean = semisup_models.subtract_1st_stage_center(pc_xyz, stage1_center)
    pred_box = semisup_models.v1_box_est(pc_xyz_submean, stage1_center, mask, one_hot_vec, end_points,
                                         is_training, vs, norm_box2D, bn_decay, c=c, scope='box_est')
    end_points['S_pred_box'] = pred_box
    end_points['S_pred_box_reg'] = tf_util.tf_convert_box_params_from_anchor_to_reg_format_multi(
        pred_box, end_points['class_ids'], end_points['dims_anchors'], end_points['orient_anchors'])
    return (logits, pred_box), end_points


def get_semi_model_final(pc, bg_pc, img, one_hot_vec, is_training, use_one_hot, vs, oracle_mask=None,
                         norm_box2D=None, bn_decay=None, c=None):
    """semisup_v1_sunrgbd.py:132-230 (model F)."""
    end_points = _base_end_points(pc, one_hot_vec)
    if not c.USE_NORMALIZED_BOX2D_AS_FEATS:
        norm_box2D = None
    with vs.variable_scope('class_agnostic'):
        logits = semisup_models.v1_inst_seg(pc, None, None, end_points, is_training, vs, bn_decay, 'inst_seg')
        if oracle_mask is not None:
            om = oracle_mask.to(pc.dtype)
            logits = torch.stack([1 - om, om], dim=2)
        mask, mean, pc_xyz, pc_xyz_stage1 = semisup_models.subtract_points_mean(pc, logits)
        stage1_center = semisup_models.v1_tnet(pc_xyz_stage1, mask, mean, None, end_points, is_training, vs,
                                               norm_box2D, bn_decay, 'tnet')
        pc_xyz_submean = semisup_models.subtract_1st_stage_center(pc_xyz, stage1_center)
        W_pred_box = semisup_models.v1_box_est(pc_xyz_submean, stage1_center, mask, None, end_points,
                                               is_training, vs, norm_box2D, bn_decay, c=c, scope='box_est')
    with vs.variable_scope('class_dependent'):
        curr_feat = end_points['feats_lv1']
        if use_one_hot:
            curr_feat = torch.cat([curr_feat, one_hot_vec], dim=1)
        output_dims = 3 + NUM_HEADING_BIN * 2 + NUM_SIZE_CLUSTER * 4
        activation_fn = leaky_relu if c.SEMI_ADV_LEAKY_RELU else torch.relu
        last_layer_fn = torch.tanh if c.SEMI_ADV_TANH_FOR_LAST_LAYER_OF_G else activation_fn
        dp = c.SEMI_ADV_DROPOUTS_FOR_G
        F_output = semisup_models.mlps_with_dropout(curr_feat, [512, 256, output_dims],
                                                    [activation_fn, last_layer_fn, None], [dp, dp, None],
                                                    is_training, vs, bn=True, bn_decay=bn_decay, c=c,
                                                    scope='box_refine')
        end_points['F_output'] = F_output
        F_pred_box = semisup_models.parse_box_output(F_output, stage1_center, end_points, 'F_')
    end_points['F_pred_box_reg'] = tf_util.tf_convert_box_params_from_anchor_to_reg_format_multi(
        F_pred_box, end_points['class_ids'], end_points['dims_anchors'], end_points['orient_anchors'])
    return (logits, W_pred_box, F_pred_box), end_points


def huber_loss(error, delta, reduce_loss=True):
    """semisup_v1_sunrgbd.py:555-564."""
    abs_error = error.abs()
    quadratic = torch.clamp(abs_error, max=delta)
    linear = abs_error - quadratic
    losses = 0.5 * quadratic ** 2 + delta * linear
    return losses.mean() if reduce_loss else losses


def get_strong_loss(pred, labels, end_points, prefix='', reg_weight=0.001, reduce_loss=True, c=None):
    """semisup_v1_sunrgbd.py:423-553: returns per-sample (mask_losses, box_losses), both (B,)."""
    Fn = torch.nn.functional
    pred_seg, pred_box = pred
    y_seg, y_center, y_orient_cls, y_orient_reg, y_dims_cls, y_dims_reg = labels
    dt = pred_seg.dtype
    NH, NS = NUM_HEADING_BIN, NUM_SIZE_CLUSTER
    B, N, _ = pred_seg.shape
    ce = Fn.cross_entropy(pred_seg.reshape(-1, 2), y_seg.reshape(-1).long(), reduction='none').reshape(B, N)
    mask_losses = ce.mean(dim=1)
    center_dist = torch.linalg.norm(y_center - end_points[prefix + 'center'], dim=-1)
    center_losses = huber_loss(center_dist, 2.0, False)
    s1_dist = torch.linalg.norm(y_center - end_points['stage1_center'], dim=-1)
    stage1_center_losses = huber_loss(s1_dist, 1.0, False)
    heading_class_losses = Fn.cross_entropy(end_points[prefix + 'heading_scores'], y_orient_cls.long(), reduction='none')
    ocls = Fn.one_hot(y_orient_cls.long(), NH).to(dt)
    hrn_label = y_orient_reg / (np.pi / NH)
    hrn_losses = huber_loss((end_points[prefix + 'heading_residuals_normalized'] * ocls).sum(dim=1) - hrn_label, 1.0, False)
    size_class_losses = Fn.cross_entropy(end_points[prefix + 'size_scores'], y_dims_cls.long(), reduction='none')
    dcls = Fn.one_hot(y_dims_cls.long(), NS).to(dt)
    dcls_t = dcls.unsqueeze(-1).repeat(1, 1, 3)
    pred_srn = (end_points[prefix + 'size_residuals_normalized'] * dcls_t).sum(dim=1)
    msa = torch.as_tensor(MEAN_DIMS_ARR, dtype=torch.float32).to(dt).unsqueeze(0)
    mean_size_label = (dcls_t * msa).sum(dim=1)
    srl_norm = y_dims_reg / mean_size_label
    srn_losses = huber_loss(torch.linalg.norm(srl_norm - pred_srn, dim=-1), 1.0, False)
    corners_3d = get_box3d_corners_sunrgbd(end_points[prefix + 'center'], end_points[prefix + 'heading_residuals'],
                                           end_points[prefix + 'size_residuals'])
    gt_mask = ocls.unsqueeze(2) * dcls.unsqueeze(1)
    corners_pred = (gt_mask.unsqueeze(-1).unsqueeze(-1) * corners_3d).sum(dim=(1, 2))
    bins = torch.as_tensor(np.arange(0, 2 * np.pi, 2 * np.pi / NH), dtype=torch.float32).to(dt)
    heading_label = (ocls * (y_orient_reg.unsqueeze(1) + bins.unsqueeze(0))).sum(dim=1)
    size_label = (dcls.unsqueeze(-1) * (msa + y_dims_reg.unsqueeze(1))).sum(dim=1)
    c_gt = get_box3d_corners_helper(y_center, heading_label, size_label)
    c_gt_flip = get_box3d_corners_helper(y_center, heading_label + np.pi, size_label)
    corners_dist = torch.minimum(torch.linalg.norm(corners_pred - c_gt, dim=-1),
                                 torch.linalg.norm(corners_pred - c_gt_flip, dim=-1))
    corners_losses = huber_loss(corners_dist, 1.0, False).mean(dim=1)
    mask_losses = c.STRONG_WEIGHT_CROSS_ENTROPY * mask_losses
    total_losses = c.STRONG_BOX_MULTIPLER * (
        c.STRONG_WEIGHT_CENTER * center_losses + c.STRONG_WEIGHT_ORIENT_CLS * heading_class_losses +
        c.STRONG_WEIGHT_DIMS_CLS * size_class_losses + c.STRONG_WEIGHT_ORIENT_REG * hrn_losses +
        c.STRONG_WEIGHT_DIMS_REG * srn_losses + c.STRONG_WEIGHT_TNET_CENTER * stage1_center_losses) + \
        c.STRONG_WEIGHT_CORNER * corners_losses
    return mask_losses, total_losses


def get_semi_loss_final(pred, labels, end_points, reduce_loss=True, c=None):
    """semisup_v1_sunrgbd.py:323-421."""
    pred_seg, W_pred_box, F_pred_box = pred
    (y_seg, y_center, y_orient_cls, y_orient_reg, y_dims_cls, y_dims_reg,
     R0_rect, P, Rtilt, K, rot_frust, box2D, img_dim, is_data_2D) = labels
    dt = pred_seg.dtype
    label = (y_seg, y_center, y_orient_cls, y_orient_reg, y_dims_cls, y_dims_reg)
    mask_losses, box_losses = get_strong_loss((pred_seg, F_pred_box), label, end_points, prefix='F_',
                                              reduce_loss=False, c=c)
    is3d = (1 - is_data_2D).to(dt)
    is2d = is_data_2D.to(dt)
    mask_loss = (mask_losses * is3d).sum() / (is3d.sum() + 1e-3)
    box_loss = (box_losses * is3d).sum() / (is3d.sum() + 1e-3)
    strong_loss = mask_loss + box_loss
    end_points['_mask_loss'], end_points['_box_loss'] = mask_loss, box_loss

    weak_loss_fns = torch.zeros((), dtype=dt)
    F_pred_box_reg = end_points['F_pred_box_reg']
    _, F_dims_reg, _ = F_pred_box_reg
    class_ids = end_points['class_ids']
    if c.WEAK_WEIGHT_INACTIVE_VOLUME != 0:
        assert len(c.WEAK_INACTIVE_VOL_LOSS_MARGINS) == 10
        margins = torch.as_tensor(c.WEAK_INACTIVE_VOL_LOSS_MARGINS, dtype=dt)
        iv = weak_losses.get_inactive_volume_loss_v1(F_dims_reg, class_ids, end_points['inactive_vol_train_classes'],
                                                     10, margins)
        weak_loss_fns = weak_loss_fns + c.WEAK_WEIGHT_INACTIVE_VOLUME * iv
    if c.WEAK_WEIGHT_INTRACLASSVAR != 0:
        icv = weak_losses.get_intraclass_variance_loss_v1(
            F_dims_reg, class_ids, end_points['intraclsdims_train_classes'], 10,
            c.WEAK_DIMS_USE_MARGIN_LOSS, c.WEAK_DIMS_SD_MARGIN, c.WEAK_DIMS_LOSS_TYPE)
        end_points['_intraclass_variance_loss'] = icv
        weak_loss_fns = weak_loss_fns + c.WEAK_WEIGHT_INTRACLASSVAR * icv
    if c.WEAK_WEIGHT_REPROJECTION != 0:
        rl = weak_losses.get_reprojection_loss(
            F_pred_box_reg, box2D, Rtilt, K, img_dim, rot_frust,
            c.WEAK_REPROJECTION_USE_SOFTMAX_PROJ, c.WEAK_REPROJECTION_SOFTMAX_SCALE,
            c.WEAK_REPROJECTION_DILATE_FACTOR, c.WEAK_REPROJECTION_CLIP_LOWERB_LOSS,
            c.WEAK_REPROJECTION_CLIP_PRED_BOX, c.WEAK_REPROJECTION_LOSS_TYPE,
            c.WEAK_TRAIN_BOX_W_REPROJECTION, reduce_loss=False, end_points=end_points)
        if c.WEAK_REPROJECTION_ONLY_ON_2D_CLS:
            weak_loss_fns = weak_loss_fns + c.WEAK_WEIGHT_REPROJECTION * (rl * is2d)
        else:
            weak_loss_fns = weak_loss_fns + c.WEAK_WEIGHT_REPROJECTION * rl
    weak_loss = weak_loss_fns.mean()
    end_points['_weak_loss'] = weak_loss
    total_loss = strong_loss + c.SEMI_MULTIPLIER_FOR_WEAK_LOSS * weak_loss
    if c.SEMI_WEIGHT_BOXPC_FIT_LOSS != 0:
        fit_losses = -torch.log(0.01 + end_points['boxpc_fit_prob'])
        if c.SEMI_BOXPC_FIT_ONLY_ON_2D_CLS:
            total_loss = total_loss + c.SEMI_WEIGHT_BOXPC_FIT_LOSS * (fit_losses * is2d).mean()
        else:
            total_loss = total_loss + c.SEMI_WEIGHT_BOXPC_FIT_LOSS * fit_losses.mean()
    if reduce_loss:
        return total_loss
    raise Exception('Not implemented')


def get_semi_loss_backbone(pred, labels, end_points, reduce_loss=True, c=None):
    """semisup_v1_sunrgbd.py:256-321 (model A), surface term :284-291 included."""
    from . import weak_losses
    pred_seg, pred_box = pred
    (y_seg, y_center, y_orient_cls, y_orient_reg, y_dims_cls, y_dims_reg, R0_rect, P, Rtilt, K, rot_frust, box2D, img_dim,
     is_data_2D) = labels
    reproj = weak_losses.get_reprojection_loss(
        end_points['S_pred_box_reg'], box2D, Rtilt, K, img_dim, rot_frust, c.WEAK_REPROJECTION_USE_SOFTMAX_PROJ,
        c.WEAK_REPROJECTION_SOFTMAX_SCALE, c.WEAK_REPROJECTION_DILATE_FACTOR, c.WEAK_REPROJECTION_CLIP_LOWERB_LOSS,
        c.WEAK_REPROJECTION_CLIP_PRED_BOX, c.WEAK_REPROJECTION_LOSS_TYPE, c.WEAK_TRAIN_BOX_W_REPROJECTION, reduce_loss=False)
    weak_loss_fns = c.WEAK_WEIGHT_REPROJECTION * reproj
    if float(c.WEAK_WEIGHT_SURFACE) != 0.0:
        surface = weak_losses.get_surface_loss(
            end_points['S_pred_box_reg'], end_points['point_cloud'][:, :, 0:3], end_points['soft_mask'], c.WEAK_SURFACE_MARGIN,
            c.WEAK_SURFACE_LOSS_SCALE_DIMS, c.WEAK_SURFACE_LOSS_WT_FOR_INNER_PTS, c.WEAK_TRAIN_SEG_W_SURFACE,
            c.WEAK_TRAIN_BOX_W_SURFACE, reduce_loss=False)
        weak_loss_fns = weak_loss_fns + c.WEAK_WEIGHT_SURFACE * surface
    mask_losses, strong_losses = get_strong_loss((pred_seg, pred_box), (y_seg, y_center, y_orient_cls, y_orient_reg, y_dims_cls,
                                                                         y_dims_reg), end_points, reduce_loss=False, c=c)
    is2d = is_data_2D.to(pred_seg.dtype)
    total_losses = (1 - is2d) * (mask_losses + strong_losses) + is2d * (weak_loss_fns * c.SEMI_MULTIPLIER_FOR_WEAK_LOSS)
    return total_losses.mean() if reduce_loss else total_losses


def get_semi_loss(pred, labels, end_points, reduce_loss=True, c=None):
    """semisup_v1_sunrgbd.py:248-254."""
    if c.SEMI_MODEL == 'A':
        return get_semi_loss_backbone(pred, labels, end_points, reduce_loss, c)
    if c.SEMI_MODEL == 'F':
        return get_semi_loss_final(pred, labels, end_points, reduce_loss, c)
    raise Exception('Not implemented SEMI_MODEL: %s' % c.SEMI_MODEL)


def convert_raw_y_box_to_reg_format(y_box, one_hot_vec):
    """semisup_v1_sunrgbd.py:584-608 (identical copy at boxpc_sunrgbd.py:206-229)."""
    Fn = torch.nn.functional
    y_centers, y_orient_cls, y_orient_reg, y_dims_cls, y_dims_reg = y_box
    dt = y_centers.dtype
    class_ids = torch.argmax(one_hot_vec, dim=1).to(torch.int32)
    dims_anchors = torch.as_tensor(MEAN_DIMS_ARR, dtype=torch.float32).to(dt)
    orient_anchors = torch.as_tensor(np.arange(0, 2 * np.pi, 2 * np.pi / NUM_HEADING_BIN), dtype=torch.float32).to(dt)
    dims_cls = Fn.one_hot(y_dims_cls.long(), NUM_SIZE_CLUSTER)
    dims_reg = tf_util.tf_expand_tile(y_dims_reg, 1, [1, NUM_SIZE_CLUSTER, 1])
    orient_cls = Fn.one_hot(y_orient_cls.long(), NUM_HEADING_BIN)
    orient_reg = tf_util.tf_expand_tile(y_orient_reg, 1, [1, NUM_HEADING_BIN])
    box = (y_centers, dims_cls, dims_reg, orient_cls, orient_reg)
    return tf_util.tf_convert_box_params_from_anchor_to_reg_format_multi(box, class_ids, dims_anchors, orient_anchors)
