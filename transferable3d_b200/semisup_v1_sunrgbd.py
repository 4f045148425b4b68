"""Mirror of sunrgbd/sunrgbd_detection/semisup_v1_sunrgbd.py model definitions
(placeholder_inputs :37-67, get_semi_model :69-79, get_semi_model_backbone :81-130,
get_semi_model_final :132-230) on the B200 path; same returned tuples and end_points keys.
get_semi_loss / get_semi_loss_final (:248-254, 323-421) and get_strong_loss (:423-553) evaluate the fused loss kernel
(csrc/loss_ops.cuh); get_iou_summary (:236-246, metrics-only tf.py_func around the missing box_util) is not built.
"""
import numpy as np
import torch

from . import runtime as rt
from . import tf_util, semisup_models
from .constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER, NUM_CLASS, MEAN_DIMS_ARR, ORIENT_ANCHORS

INPUT_IMG_CHANNELS = 3


def placeholder_inputs(batch_size, num_point, num_channel, device='cuda'):
    """semisup_v1_sunrgbd.py:37-67: the 18 inputs, in the reference order, as zero tensors of the
    placeholder shape/dtype (img has no static H,W -> (B,0,0,3))."""
    f, i = torch.float32, torch.int32
    B, N, C = batch_size, num_point, num_channel
    Z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=device)
    return (Z((B, N, C), f), Z((B, N, C), f), Z((B, 0, 0, INPUT_IMG_CHANNELS), f), Z((B, NUM_CLASS), f),
            Z((B, N), i), Z((B, 3), f), Z((B,), i), Z((B,), f), Z((B,), i), Z((B, 3), f),
            Z((B, 3, 3), f), Z((B, 3, 4), f), Z((B, 3, 3), f), Z((B, 3, 3), f),
            Z((B, 1), f), Z((B, 4), f), Z((B, 2), f), Z((B,), i))


def _base_end_points(pc, one_hot_vec):
    st = rt.store()
    return {'point_cloud': pc, 'class_one_hot': one_hot_vec,
            'class_ids': torch.argmax(one_hot_vec, dim=1).to(torch.int32),
            'dims_anchors': st.const('MEAN_DIMS_ARR', MEAN_DIMS_ARR),
            'orient_anchors': st.const('ORIENT_ANCHORS', ORIENT_ANCHORS)}


def get_semi_model(pc, bg_pc, img, one_hot_vec, is_training, use_one_hot, oracle_mask=None, norm_box2D=None,
                   bn_decay=None, c=None):
    """semisup_v1_sunrgbd.py:69-79."""
    if c.SEMI_MODEL == 'A':
        return get_semi_model_backbone(pc, bg_pc, img, one_hot_vec, is_training, use_one_hot, oracle_mask=oracle_mask,
                                       norm_box2D=norm_box2D, bn_decay=bn_decay, c=c)
    elif c.SEMI_MODEL == 'F':
        return get_semi_model_final(pc, bg_pc, img, one_hot_vec, is_training, use_one_hot, oracle_mask=oracle_mask,
                                    norm_box2D=norm_box2D, bn_decay=bn_decay, c=c)
    else:
        raise Exception('Not implemented SEMI_MODEL: %s' % c.SEMI_MODEL)


def get_semi_model_backbone(pc, bg_pc, img, one_hot_vec, is_training, use_one_hot, oracle_mask=None, norm_box2D=None,
                            bn_decay=None, c=None):
    """semisup_v1_sunrgbd.py:81-130 (model A)."""
    if oracle_mask is not None:
        raise NotImplementedError          # semisup_v1_sunrgbd.py:96: model A takes no oracle mask
    end_points = _base_end_points(pc, one_hot_vec)
    img_feats = None
    if not use_one_hot:
        one_hot_vec = None
    if not c.USE_NORMALIZED_BOX2D_AS_FEATS:
        norm_box2D = None
    logits = semisup_models.v1_inst_seg(pc, img_feats, one_hot_vec, end_points, is_training, bn_decay=bn_decay,
                                        scope='inst_seg')
    end_points['soft_mask'] = torch.softmax(logits, dim=-1)[:, :, 1]
    mask, mask_xyz_mean, pc_xyz, pc_xyz_stage1 = semisup_models.subtract_points_mean(pc, logits, scope='subtract_points_mean')
    stage1_center = semisup_models.v1_tnet(pc_xyz_stage1, mask, mask_xyz_mean, one_hot_vec, end_points, is_training,
                                           norm_box2D=norm_box2D, bn_decay=bn_decay, scope='tnet')
    pc_xyz_submean = semisup_models.subtract_1st_stage_center(pc_xyz, stage1_center, scope='subtract_tnet_center')
    pred_box = semisup_models.v1_box_est(pc_xyz_submean, stage1_center, mask, one_hot_vec, end_points, is_training,
                                         norm_box2D=norm_box2D, bn_decay=bn_decay, c=c, scope='box_est')
    end_points['S_pred_box'] = pred_box
    end_points['S_pred_box_reg'] = end_points.pop('_box_reg_fused')      # anchor->reg fused into the parse kernel
    pred = (logits, pred_box)
    return pred, end_points


def get_semi_model_final(pc, bg_pc, img, one_hot_vec, is_training, use_one_hot, oracle_mask=None, norm_box2D=None,
                         bn_decay=None, c=None):
    """semisup_v1_sunrgbd.py:132-230 (model F)."""
    end_points = _base_end_points(pc, one_hot_vec)
    img_feats = None
    if not c.USE_NORMALIZED_BOX2D_AS_FEATS:
        norm_box2D = None
    with rt.variable_scope('class_agnostic'):
        logits = semisup_models.v1_inst_seg(pc, img_feats, None, end_points, is_training, bn_decay=bn_decay,
                                            scope='inst_seg')
        if oracle_mask is not None:
            om = oracle_mask.to(torch.float32)
            logits = torch.stack([1 - om, om], dim=2).contiguous()
        mask, mask_xyz_mean, pc_xyz, pc_xyz_stage1 = semisup_models.subtract_points_mean(pc, logits,
                                                                                       scope='subtract_points_mean')
        end_points['_mask'] = mask
        stage1_center = semisup_models.v1_tnet(pc_xyz_stage1, mask, mask_xyz_mean, None, end_points, is_training,
                                               norm_box2D=norm_box2D, bn_decay=bn_decay, scope='tnet')
        pc_xyz_submean = semisup_models.subtract_1st_stage_center(pc_xyz, stage1_center, scope='subtract_tnet_center')
        W_pred_box = semisup_models.v1_box_est(pc_xyz_submean, stage1_center, mask, None, end_points, is_training,
                                               norm_box2D=norm_box2D, bn_decay=bn_decay, c=c, scope='box_est')
        end_points.pop('_box_reg_fused')
    with rt.variable_scope('class_dependent'):
        curr_feat = end_points['feats_lv1']
        if use_one_hot:
            curr_feat = torch.cat([curr_feat, rt.f32(one_hot_vec)], dim=1).contiguous()
        output_dims = 3 + NUM_HEADING_BIN * 2 + NUM_SIZE_CLUSTER * 4
        activation_fn = 'leaky_relu' if c.SEMI_ADV_LEAKY_RELU else 'relu'
        last_layer_fn = 'tanh' if c.SEMI_ADV_TANH_FOR_LAST_LAYER_OF_G else activation_fn
        dropout = c.SEMI_ADV_DROPOUTS_FOR_G
        F_output = semisup_models.mlps_with_dropout(curr_feat, layers=[512, 256, output_dims],
                                                    activation_fns=[activation_fn, last_layer_fn, None],
                                                    keep_probs=[dropout, dropout, None], is_training=is_training,
                                                    bn=True, bn_decay=bn_decay, c=c, scope='box_refine', reuse=None)
        end_points['F_output'] = F_output
        F_pred_box, F_reg = semisup_models.parse_into_end_points(F_output, stage1_center, end_points, 'F_')
    end_points['F_pred_box_reg'] = F_reg
    pred = (logits, W_pred_box, F_pred_box)
    return pred, end_points


# ------------------------------------------------------------------------------------------ losses
def _label_feed(labels):
    (y_seg, y_center, y_orient_cls, y_orient_reg, y_dims_cls, y_dims_reg, R0_rect, P, Rtilt, K, rot_frust, box2D, img_dim,
     is_data_2D) = labels
    return dict(labels=y_seg, centers=y_center, y_orient_cls=y_orient_cls, y_orient_reg=y_orient_reg, y_dims_cls=y_dims_cls,
                y_dims_reg=y_dims_reg, Rtilt=Rtilt, K=K, rot_frust=rot_frust, box2D=box2D, img_dim=img_dim, is_data_2D=is_data_2D)


def _fit_logits(end_points):
    if 'boxpc_fit_logits' in end_points:
        return rt.f32(end_points['boxpc_fit_logits'])
    p = rt.f32(end_points['boxpc_fit_prob'])
    return torch.stack([torch.log1p(-p), torch.log(p)], dim=1).contiguous()       # softmax of these is (1-p, p)


def get_semi_loss_final(pred, labels, end_points, reduce_loss=True, c=None):
    """semisup_v1_sunrgbd.py:323-421 -> scalar total loss; the terms land in end_points['semi_loss_terms'] =
    [total, mask_loss, box_loss, intraclass_var, weak_loss, fit_loss, reproj mean] and the gradients w.r.t. F_output /
    stage1_center / fit logits in end_points['semi_loss_grads']."""
    from . import losses
    if not reduce_loss:
        raise Exception('Not implemented')
    pred_seg = rt.f32(pred[0])
    F_output = rt.f32(end_points['F_output'])
    fit = _fit_logits(end_points) if c.SEMI_WEIGHT_BOXPC_FIT_LOSS != 0 else None
    res = losses.semi_loss(c, F_output, rt.f32(end_points['stage1_center']), rt.f32(end_points['class_one_hot']), _label_feed(labels),
                           F_output.device, logits=pred_seg, fit_logits=fit)
    end_points['semi_loss_terms'] = res['total']
    end_points['semi_loss_grads'] = {'F_output': res['dF'], 'stage1_center': res['ds1'], 'boxpc_fit_logits': res['dfit']}
    end_points['reproj_loss'] = res['per_sample'][:, 2]
    return res['total'][0]


def get_iou_summary(pred_box, y_box, end_points, name_prefix=''):
    """semisup_v1_sunrgbd.py:236-246: the tf.py_func around roi_seg_box3d_dataset.compute_box3d_iou, here one GPU kernel
    (box_util.compute_box3d_iou); sets end_points[name_prefix + 'iou2ds' / 'iou3ds'] (B,).
    pred_box = (center, dims_cls scores, dims_reg (B,NS,3), orient_cls scores, orient_reg (B,NH));
    y_box = (center, dims_cls, dims_reg, orient_cls, orient_reg) labels."""
    from . import box_util
    pred_center, pred_dims_cls, pred_dims_reg, pred_orient_cls, pred_orient_reg = pred_box
    y_center, y_dims_cls, y_dims_reg, y_orient_cls, y_orient_reg = y_box
    iou2ds, iou3ds = box_util.compute_box3d_iou(pred_center, pred_orient_cls, pred_orient_reg, pred_dims_cls, pred_dims_reg,
                                                y_center, y_orient_cls, y_orient_reg, y_dims_cls, y_dims_reg)
    end_points[name_prefix + 'iou2ds'] = iou2ds
    end_points[name_prefix + 'iou3ds'] = iou3ds
    return iou2ds, iou3ds


def get_semi_loss_backbone(pred, labels, end_points, reduce_loss=True, c=None):
    """semisup_v1_sunrgbd.py:256-321 (model A): mean_B[(1 - is2D) * (mask + strong) + is2D * (W_r * reprojection +
    W_s * surface) * SEMI_MULTIPLIER_FOR_WEAK_LOSS].  Forward values from the fused loss kernel (two launches: strong terms
    on the parsed head output, reprojection on S_pred_box_reg) and, when WEAK_WEIGHT_SURFACE != 0, the surface-loss kernel on
    (S_pred_box_reg, point_cloud xyz, soft_mask) (:284-291).  The metrics-only get_iou_summary of :316 is available
    separately (get_iou_summary)."""
    from . import weak_losses
    (y_seg, y_center, y_orient_cls, y_orient_reg, y_dims_cls, y_dims_reg, R0_rect, P, Rtilt, K, rot_frust, box2D, img_dim,
     is_data_2D) = labels
    reproj = weak_losses.get_reprojection_loss(
        end_points['S_pred_box_reg'], box2D, Rtilt, K, img_dim, rot_frust,
        use_softmax_projection=c.WEAK_REPROJECTION_USE_SOFTMAX_PROJ, softmax_scale_factor=c.WEAK_REPROJECTION_SOFTMAX_SCALE,
        dilate_factor=c.WEAK_REPROJECTION_DILATE_FACTOR, clip_lower_b_loss=c.WEAK_REPROJECTION_CLIP_LOWERB_LOSS,
        clip_pred_box=c.WEAK_REPROJECTION_CLIP_PRED_BOX, loss_type=c.WEAK_REPROJECTION_LOSS_TYPE,
        train_box=c.WEAK_TRAIN_BOX_W_REPROJECTION, end_points=end_points, reduce_loss=False, scope='reprojection_loss')
    mask_losses, strong_losses = get_strong_loss(pred, (y_seg, y_center, y_orient_cls, y_orient_reg, y_dims_cls, y_dims_reg),
                                                 end_points, reduce_loss=False, c=c)
    is2d = is_data_2D.to(torch.float32)
    weak = float(c.WEAK_WEIGHT_REPROJECTION) * reproj
    if float(c.WEAK_WEIGHT_SURFACE) != 0.0:
        surface = weak_losses.get_surface_loss(
            end_points['S_pred_box_reg'], end_points['point_cloud'][:, :, 0:3].contiguous(), end_points['soft_mask'],
            margin=c.WEAK_SURFACE_MARGIN, scale_dims_factor=c.WEAK_SURFACE_LOSS_SCALE_DIMS,
            weight_for_points_within=c.WEAK_SURFACE_LOSS_WT_FOR_INNER_PTS, train_seg=c.WEAK_TRAIN_SEG_W_SURFACE,
            train_box=c.WEAK_TRAIN_BOX_W_SURFACE, end_points=end_points, reduce_loss=False, scope='surface_loss')
        weak = weak + float(c.WEAK_WEIGHT_SURFACE) * surface
    total_losses = (1.0 - is2d) * (mask_losses + strong_losses) + is2d * (weak * float(c.SEMI_MULTIPLIER_FOR_WEAK_LOSS))
    return total_losses.mean() if reduce_loss else total_losses


def get_semi_loss(pred, labels, end_points, reduce_loss=True, c=None):
    """semisup_v1_sunrgbd.py:248-254."""
    if c.SEMI_MODEL == 'A':
        return get_semi_loss_backbone(pred, labels, end_points, reduce_loss, c)
    if c.SEMI_MODEL == 'F':
        return get_semi_loss_final(pred, labels, end_points, reduce_loss, c)
    raise Exception('Not implemented SEMI_MODEL: %s' % c.SEMI_MODEL)


def get_strong_loss(pred, labels, end_points, prefix='', reg_weight=0.001, reduce_loss=True, c=None):
    """semisup_v1_sunrgbd.py:423-553 -> per-sample (mask_losses, box_losses), both (B,) (their means if reduce_loss).
    The head output is read from end_points['F_output'] (prefix 'F_') or end_points['box_params'] (prefix '').
    `reg_weight` keeps the reference's positional signature; the reference body never reads it either."""
    from . import losses
    from .config import cfg as _cfg
    pred_seg = rt.f32(pred[0])
    out = rt.f32(end_points['F_output' if prefix == 'F_' else prefix + 'box_params'])
    y_seg, y_center, y_orient_cls, y_orient_reg, y_dims_cls, y_dims_reg = labels
    B, dev = out.shape[0], out.device
    z = lambda *s_, dt=torch.float32: torch.zeros(s_, dtype=dt, device=dev)
    feed = dict(labels=y_seg, centers=y_center, y_orient_cls=y_orient_cls, y_orient_reg=y_orient_reg, y_dims_cls=y_dims_cls,
                y_dims_reg=y_dims_reg, Rtilt=z(B, 3, 3), K=z(B, 3, 3), rot_frust=z(B), box2D=z(B, 4), img_dim=z(B, 2),
                is_data_2D=z(B, dt=torch.int32))
    cc = _cfg(**{k: getattr(c, k) for k in ('STRONG_WEIGHT_CROSS_ENTROPY', 'STRONG_BOX_MULTIPLER', 'STRONG_WEIGHT_CENTER',
                                             'STRONG_WEIGHT_ORIENT_CLS', 'STRONG_WEIGHT_ORIENT_REG', 'STRONG_WEIGHT_DIMS_CLS',
                                             'STRONG_WEIGHT_DIMS_REG', 'STRONG_WEIGHT_TNET_CENTER', 'STRONG_WEIGHT_CORNER')},
              WEAK_WEIGHT_REPROJECTION=0., WEAK_WEIGHT_INTRACLASSVAR=0., SEMI_WEIGHT_BOXPC_FIT_LOSS=0.)
    res = losses.semi_loss(cc, out, rt.f32(end_points['stage1_center']), rt.f32(end_points['class_one_hot']), feed, dev, logits=pred_seg)
    mask_losses = res['per_sample'][:, 5].contiguous()
    box_losses = res['per_sample'][:, 0].contiguous()
    if reduce_loss:
        return mask_losses.mean(), box_losses.mean()
    return mask_losses, box_losses


def convert_raw_y_box_to_reg_format(y_box, one_hot_vec):
    """semisup_v1_sunrgbd.py:588-607: the reference defines this twice, verbatim (boxpc_sunrgbd.py:208-229); one implementation here."""
    from . import boxpc_sunrgbd
    return boxpc_sunrgbd.convert_raw_y_box_to_reg_format(y_box, one_hot_vec)
