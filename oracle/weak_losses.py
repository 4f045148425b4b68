"""Weak losses: restatement of models/weak_losses.py:15-36 (range deviation), :38-67 (inactive
volume), :69-238 (relaxed reprojection), :267-291 (intra-class variance).
tf.losses.huber_loss(labels, predictions, delta=1, reduction=NONE) is the element-wise huber;
its default reduction SUM_BY_NONZERO_WEIGHTS = sum / #elements with safe-div (0 when empty).
"""
import torch

from . import tf_util


def _tf_huber(labels, predictions, delta=1.0):
    err = (predictions - labels).abs()
    quad = torch.clamp(err, max=delta)
    lin = err - quad
    return 0.5 * quad * quad + delta * lin


def _tf_mse(labels, predictions):
    return (predictions - labels) ** 2


def loss_for_deviation_from_range(val, lower_b, upper_b, loss='huber'):
    """weak_losses.py:15-36."""
    viol_lo = (val < lower_b).to(val.dtype)
    viol_hi = (val > upper_b).to(val.dtype)
    if loss == 'huber':
        lo, hi = _tf_huber(lower_b, val), _tf_huber(upper_b, val)
    elif loss == 'mse':
        lo, hi = _tf_mse(lower_b, val), _tf_mse(upper_b, val)
    else:
        raise Exception('Not implemented: %s' % loss)
    return viol_lo * lo + viol_hi * hi


def get_inactive_volume_loss_v1(dims_reg, y_class, inactive_vol_train_classes, num_classes, inactive_vol_loss_margins):
    """weak_losses.py:38-67 (mean of an empty group is NaN -> 0)."""
    assert inactive_vol_loss_margins.shape[0] == num_classes == len(inactive_vol_train_classes)
    group_losses = []
    for i, train_on_cls in enumerate(inactive_vol_train_classes):
        if not train_on_cls:
            continue
        g = dims_reg[y_class.long() == i]
        if g.shape[0] == 0:
            group_losses.append(torch.zeros((), dtype=dims_reg.dtype))
            continue
        vol = g.prod(dim=1)
        group_losses.append(torch.clamp(inactive_vol_loss_margins[i] - vol, min=0.).mean())
    return torch.stack(group_losses).mean()


def get_surface_loss(pred_box_reg, pc_xyz, soft_mask, margin, scale_dims_factor, weight_for_points_within, train_seg, train_box,
                     reduce_loss=True):
    """weak_losses.py:240-265.  `mask` (the stop-gradient copy) is computed and not used by the reference: the product is
    taken with soft_mask itself (:257), so the mask gradient flows whatever train_seg says; weight_for_points_within only
    appears in the commented-out variant.  Both replicated."""
    center_reg, dims_reg, orient_reg = pred_box_reg
    center_reg = center_reg if train_box[0] else center_reg.detach()
    dims_reg = dims_reg if train_box[1] else dims_reg.detach()
    orient_reg = orient_reg if train_box[2] else orient_reg.detach()
    box = (center_reg, dims_reg * scale_dims_factor, orient_reg)
    d = tf_util.tf_distance_to_closest_3D_box_surface_multi(pc_xyz, box)       # (B,N)
    loss = torch.clamp(d - margin, min=0.) * soft_mask
    loss = loss.mean(dim=1)
    return loss.mean() if reduce_loss else loss


def get_reprojection_loss(pred_box_reg, box2D, Rtilts, Ks, img_dims, rot_frust, use_softmax_projection,
                          softmax_scale_factor, dilate_factor, clip_lower_b_loss, clip_pred_box, loss_type,
                          train_box, reduce_loss=True, scope=None, end_points=None):
    """weak_losses.py:69-238."""
    center_reg, dims_reg, orient_reg = pred_box_reg
    center_reg = center_reg if train_box[0] else center_reg.detach()
    dims_reg = dims_reg if train_box[1] else dims_reg.detach()
    orient_reg = orient_reg if train_box[2] else orient_reg.detach()
    corrected = tf_util.tf_rot_box_params_multi((center_reg, dims_reg, orient_reg), 1 * rot_frust)
    _, pts = tf_util.tf_create_3D_box_by_vertices_multi(corrected, apply_translation=True)
    if use_softmax_projection:
        pbox = tf_util.tf_get_2D_bbox_of_softmax_projection_sunrgbd_multi(pts, Rtilts, Ks, softmax_scale_factor)
    else:
        pbox = tf_util.tf_get_2D_bbox_of_projection_sunrgbd_multi(pts, Rtilts, Ks)
    dev = loss_for_deviation_from_range
    if clip_pred_box:
        pbox = tf_util.tf_clip_2D_bbox_to_image_dims_multi(pbox, img_dims)
        small = tf_util.tf_clip_2D_bbox_to_image_dims_multi(box2D, img_dims)
        big = tf_util.tf_clip_2D_bbox_to_image_dims_multi(tf_util.tf_dilate_2D_bboxes(box2D, dilate_factor), img_dims)
        left = dev(pbox[:, 0], big[:, 0], small[:, 0], loss_type)
        top = dev(pbox[:, 1], big[:, 1], small[:, 1], loss_type)
        right = dev(pbox[:, 2], small[:, 2], big[:, 2], loss_type)
        bot = dev(pbox[:, 3], small[:, 3], big[:, 3], loss_type)
        loss = left + top + right + bot
    else:
        small = tf_util.tf_clip_2D_bbox_to_image_dims_multi(box2D, img_dims)
        big = tf_util.tf_dilate_2D_bboxes(box2D, dilate_factor)
        big_clip = tf_util.tf_clip_2D_bbox_to_image_dims_multi(big, img_dims)
        not_clipped = (big == big_clip).to(box2D.dtype)
        if clip_lower_b_loss:
            left = not_clipped[:, 0] * dev(pbox[:, 0], big_clip[:, 0], small[:, 0], loss_type)
            top = not_clipped[:, 1] * dev(pbox[:, 1], big_clip[:, 1], small[:, 1], loss_type)
            right = not_clipped[:, 2] * dev(pbox[:, 2], small[:, 2], big_clip[:, 2], loss_type)
            bot = not_clipped[:, 3] * dev(pbox[:, 3], small[:, 3], big_clip[:, 3], loss_type)
            loss = torch.clamp(left + top + right + bot, max=1000.)
        else:
            lf = _tf_huber if loss_type == 'huber' else _tf_mse
            z = torch.zeros_like(pbox[:, 0])

            def side(k, inner_ok_if_less):
                inner = lf(small[:, k], pbox[:, k])
                outer = lf(big_clip[:, k], pbox[:, k])
                if inner_ok_if_less:        # left / top
                    inner = torch.where(pbox[:, k] < small[:, k], z, inner)
                    outer = torch.where(pbox[:, k] > big_clip[:, k], z, outer)
                else:                       # right / bottom
                    inner = torch.where(pbox[:, k] > small[:, k], z, inner)
                    outer = torch.where(pbox[:, k] < big_clip[:, k], z, outer)
                return inner + outer * not_clipped[:, k]
            left, top, right, bot = side(0, True), side(1, True), side(2, False), side(3, False)
            loss = (torch.clamp(left, max=1000.) + torch.clamp(top, max=1000.) +
                    torch.clamp(right, max=1000.) + torch.clamp(bot, max=1000.))
    if reduce_loss:
        loss = loss.mean()
    if end_points is not None:
        end_points['reproj_pred_box3D_pts'] = pts
        end_points['reproj_proj_pred_box'] = pbox
        end_points['reproj_gt_box2D_lowerb'] = small
        end_points['reproj_gt_box2D_upperb'] = big
        end_points['left_loss'], end_points['top_loss'] = left, top
        end_points['right_loss'], end_points['bot_loss'] = right, bot
        end_points['reproj_loss'] = loss
    return loss


def get_intraclass_variance_loss_v1(dims_reg, y_class, intraclsdims_train_classes, num_classes, use_margin_loss,
                                    dims_sd_margin, loss_type, scope=None):
    """weak_losses.py:267-291. use_margin_loss / dims_sd_margin are ignored by the reference.
    Per trained class: tf.losses.{huber,mse}(labels=stop_grad(group mean), predictions=group)
    with the default reduction (sum / #elements, 0 for an empty group); mean over trained
    classes, empty ones included as 0."""
    group_losses = []
    for i, train_on_cls in enumerate(intraclsdims_train_classes):
        if not train_on_cls:
            continue
        g = dims_reg[y_class.long() == i]
        if g.shape[0] == 0:
            group_losses.append(torch.zeros((), dtype=dims_reg.dtype))
            continue
        mean = g.mean(dim=0, keepdim=True).detach().expand_as(g)
        l = _tf_huber(mean, g) if loss_type == 'huber' else _tf_mse(mean, g)
        group_losses.append(l.sum() / l.numel())
    return torch.stack(group_losses).mean()
