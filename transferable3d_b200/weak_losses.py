"""Mirror of models/weak_losses.py on the B200 path: get_reprojection_loss (:69-238) and
get_intraclass_variance_loss_v1 (:267-291) with the reference's argument lists, evaluated by the fused loss kernel
(csrc/loss_ops.cuh, t3d_semi_loss) with every other term switched off.  Forward values; the gradients of these terms
inside a training step are produced by the same kernel through train_semisup_adv.SemiAdvTrainGraph.
get_surface_loss (:240-265) has its own kernel (csrc/surface_ops.cuh, t3d_surface_loss: forward and backward in one pass
over the points); get_inactive_volume_loss_v1 (:38-67) is a one-CTA kernel (t3d_inactive_volume_loss).  get_D_loss /
get_G_loss belong to the dead discriminator branches and are not built."""
import numpy as np
import torch

from . import losses
from .config import cfg
from .constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER, NUM_CLASS

_ZERO = dict(STRONG_WEIGHT_CROSS_ENTROPY=0., STRONG_BOX_MULTIPLER=0., STRONG_WEIGHT_CORNER=0., WEAK_WEIGHT_INTRACLASSVAR=0.,
             WEAK_WEIGHT_REPROJECTION=0., SEMI_WEIGHT_BOXPC_FIT_LOSS=0., SEMI_MULTIPLIER_FOR_WEAK_LOSS=1.)


def _dummy(B, dev):
    z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=dev)
    feed = dict(centers=z(B, 3), y_orient_cls=z(B, dt=torch.int32), y_orient_reg=z(B), y_dims_cls=z(B, dt=torch.int32), y_dims_reg=z(B, 3),
                Rtilt=z(B, 3, 3), K=z(B, 3, 3), rot_frust=z(B), box2D=z(B, 4), img_dim=z(B, 2), is_data_2D=z(B, dt=torch.int32))
    oh = z(B, NUM_CLASS)
    oh[:, 0] = 1
    return feed, z(B, 3 + 2 * NUM_HEADING_BIN + 4 * NUM_SIZE_CLUSTER), z(B, 3), oh


def get_reprojection_loss(pred_box_reg, box2D, Rtilts, Ks, img_dims, rot_frust, use_softmax_projection, softmax_scale_factor,
                          dilate_factor, clip_lower_b_loss, clip_pred_box, loss_type, train_box, reduce_loss=True, scope=None,
                          end_points=None):
    """weak_losses.py:69-238: pred_box_reg = (center (B,3), dims (B,3), orient (B,)) -> (B,) losses (scalar mean if reduce_loss)."""
    center, dims, orient = pred_box_reg
    dev, B = center.device, center.shape[0]
    feed, out0, s0, oh = _dummy(B, dev)
    feed.update(Rtilt=Rtilts, K=Ks, rot_frust=rot_frust, box2D=box2D, img_dim=img_dims)
    c = cfg(WEAK_REPROJECTION_USE_SOFTMAX_PROJ=bool(use_softmax_projection), WEAK_REPROJECTION_SOFTMAX_SCALE=float(softmax_scale_factor),
            WEAK_REPROJECTION_DILATE_FACTOR=float(dilate_factor), WEAK_REPROJECTION_CLIP_LOWERB_LOSS=bool(clip_lower_b_loss),
            WEAK_REPROJECTION_CLIP_PRED_BOX=bool(clip_pred_box), WEAK_REPROJECTION_LOSS_TYPE=loss_type,
            WEAK_TRAIN_BOX_W_REPROJECTION=list(train_box), WEAK_REPROJECTION_ONLY_ON_2D_CLS=False,
            **dict(_ZERO, WEAK_WEIGHT_REPROJECTION=1.))
    reg = torch.cat([center, dims, orient.reshape(B, 1)], dim=1).to(torch.float32).contiguous()
    res = losses.semi_loss(c, out0, s0, oh, feed, dev, reg_in=reg, finish=False)
    loss = res['per_sample'][:, 2].contiguous()
    if end_points is not None:
        end_points['reproj_loss'] = loss
        end_points['reproj_grad_box_reg'] = res['g_reg'] * B          # d sum_b loss_b / d (center, dims, orient)
    return loss.mean() if reduce_loss else loss


def get_intraclass_variance_loss_v1(dims_reg, y_class, intraclsdims_train_classes, num_classes, use_margin_loss, dims_sd_margin,
                                    loss_type, scope=None):
    """weak_losses.py:267-291 (use_margin_loss / dims_sd_margin are ignored by the reference) -> scalar."""
    dev, B = dims_reg.device, dims_reg.shape[0]
    assert num_classes == NUM_CLASS == len(intraclsdims_train_classes)
    feed, out0, s0, _ = _dummy(B, dev)
    oh = torch.nn.functional.one_hot(y_class.long(), NUM_CLASS).to(torch.float32).contiguous()
    c = cfg(WEAK_DIMS_LOSS_TYPE=loss_type, **dict(_ZERO, WEAK_WEIGHT_INTRACLASSVAR=1.))
    mask = sum(1 << i for i, t in enumerate(intraclsdims_train_classes) if t)
    reg = torch.cat([torch.zeros((B, 3), device=dev), dims_reg.to(torch.float32), torch.zeros((B, 1), device=dev)], dim=1).contiguous()
    res = losses.semi_loss(c, out0, s0, oh, feed, dev, reg_in=reg, icv_mask=mask, finish=False)
    return res['total'][3]


def get_surface_loss(pred_box_reg, pc_xyz, soft_mask, margin, scale_dims_factor, weight_for_points_within, train_seg, train_box,
                     reduce_loss=True, scope=None, end_points=None, upstream=None):
    """weak_losses.py:240-265: pred_box_reg = (center (B,3), dims (B,3), orient (B,)), pc_xyz (B,N,>=3), soft_mask (B,N) ->
    (B,) losses (scalar mean if reduce_loss).  weight_for_points_within and train_seg do not enter the reference's result
    (see csrc/surface_ops.cuh).  With `upstream` (B,) = d total / d loss_b the same launch also returns the gradients:
    end_points['surface_grad_box_reg'] (B,7) w.r.t. (center, dims, orient), gated by train_box, and
    end_points['surface_grad_soft_mask'] (B,N)."""
    import ctypes
    from . import runtime as rt
    from ._lib import t3d_surface_loss_args, load, check, ptr, stream
    center, dims, orient = (rt.f32(t) for t in pred_box_reg)
    pc, sm = rt.f32(pc_xyz), rt.f32(soft_mask)
    B, N, C = pc.shape
    dev = pc.device
    loss = torch.empty((B,), dtype=torch.float32, device=dev)
    up = rt.f32(upstream) if upstream is not None else None
    g_box = torch.zeros((B, 7), dtype=torch.float32, device=dev) if up is not None else None
    g_mask = torch.empty((B, N), dtype=torch.float32, device=dev) if up is not None else None
    a = t3d_surface_loss_args(ptr(pc), C, ptr(sm), ptr(center), ptr(dims), ptr(orient.reshape(B).contiguous()), B, N, float(margin),
                              float(scale_dims_factor), int(bool(train_box[0])), int(bool(train_box[1])), int(bool(train_box[2])),
                              ptr(up), ptr(loss), ptr(g_box), ptr(g_mask))
    check(load().t3d_surface_loss(ctypes.byref(a), stream()))
    if end_points is not None:
        end_points['surface_loss'] = loss
        if up is not None:
            end_points['surface_grad_box_reg'] = g_box
            end_points['surface_grad_soft_mask'] = g_mask
    return loss.mean() if reduce_loss else loss


def get_inactive_volume_loss_v1(dims_reg, y_class, inactive_vol_train_classes, num_classes, inactive_vol_loss_margins, scope=None):
    """weak_losses.py:38-67: dims_reg (B,3), y_class (B,), train flags and margins per class -> scalar."""
    from . import runtime as rt
    from ._lib import call, ptr, stream
    dims = rt.f32(dims_reg)
    dev, B = dims.device, dims.shape[0]
    margins = rt.f32(torch.as_tensor(np.asarray(inactive_vol_loss_margins.cpu() if torch.is_tensor(inactive_vol_loss_margins)
                                                else inactive_vol_loss_margins, dtype=np.float32)).to(dev))
    assert margins.shape[0] == num_classes == len(inactive_vol_train_classes)
    oh = torch.nn.functional.one_hot(y_class.long(), num_classes).to(torch.float32).contiguous()
    mask = sum(1 << i for i, t in enumerate(inactive_vol_train_classes) if t)
    out = torch.empty((1,), dtype=torch.float32, device=dev)
    call('t3d_inactive_volume_loss', ptr(dims), ptr(oh), ptr(margins), B, int(num_classes), mask, 0.0, 0.0, ptr(out), None, None, stream())
    return out[0]
