"""Host side of the fused semi-supervised loss kernel (csrc/loss_ops.cuh): semisup_v1_sunrgbd.get_semi_loss_final
(semisup_v1_sunrgbd.py:323-421) = get_strong_loss(prefix 'F_') (:423-553) + weak_losses.get_reprojection_loss
(weak_losses.py:69-238) + get_intraclass_variance_loss_v1 (:267-291) + the BoxPC fit loss, value and gradients.
Used by train_semisup_adv.SemiAdvTrainGraph and by the reference-named wrappers in semisup_v1_sunrgbd / weak_losses."""
import ctypes

import numpy as np
import torch

from . import tf_util
from ._lib import ptr, stream, call, t3d_semi_loss_args
from .constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER, NUM_CLASS, MEAN_DIMS_ARR

ALL_CLASSES = ['bed', 'table', 'sofa', 'chair', 'toilet', 'desk', 'dresser', 'night_stand', 'bookshelf', 'bathtub']
LABEL_KEYS = ('centers', 'y_orient_cls', 'y_orient_reg', 'y_dims_cls', 'y_dims_reg', 'Rtilt', 'K', 'rot_frust', 'box2D', 'img_dim',
              'is_data_2D')


def icv_train_mask(FLAGS):
    """intraclsdims_train_classes of train_semisup_adv.py:320-321 as a bit mask over ALL_CLASSES."""
    test_cls = getattr(FLAGS, 'TEST_CLS', None) or []
    icv = [(cls in test_cls) for cls in ALL_CLASSES] if FLAGS.SEMI_INTRACLSDIMS_ONLY_ON_2D_CLS else [True] * len(ALL_CLASSES)
    return sum(1 << i for i, t in enumerate(icv) if t)


def inactive_vol_train_mask(FLAGS):
    """inactive_vol_train_classes of train_semisup_adv.py:322-323 as a bit mask over ALL_CLASSES."""
    test_cls = getattr(FLAGS, 'TEST_CLS', None) or []
    iv = [(cls in test_cls) for cls in ALL_CLASSES] if FLAGS.WEAK_INACTIVE_VOL_ONLY_ON_2D_CLS else [True] * len(ALL_CLASSES)
    return sum(1 << i for i, t in enumerate(iv) if t)


def _consts(dev):
    D = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32)).to(dev).contiguous()
    return D(MEAN_DIMS_ARR), D(np.arange(0, 2 * np.pi, 2 * np.pi / NUM_HEADING_BIN))


def semi_loss(FLAGS, F_output, stage1_center, one_hot, feed, dev, logits=None, mask_losses=None, fit_logits=None, F_reg=None,
              icv_mask=None, finish=True, mean_size=None, orient_anchors=None, reg_in=None, model_a=False):
    """Runs t3d_seg_ce (if `logits` is given), t3d_class_dims_stats and t3d_semi_loss.
    feed: dict with LABEL_KEYS (+ 'labels' when logits is given).  finish=True also folds g_reg into dF / ds1
    (t3d_box_reg_backward); the training graph passes finish=False, adds the BoxPC input gradient to g_reg and calls
    finish_box_reg itself.  model_a: the mixing of get_semi_loss_backbone (semisup_v1_sunrgbd.py:256-321) --
    mean_B[(1 - is2D)(mask + strong) + is2D * mult * W_r * reprojection]: strong terms weighted 1 / B, reprojection on the 2D
    samples only, no intra-class-variance / fit / inactive-volume terms (the surface term is added by the caller).
    Returns dict(total[8], dF, ds1, g_reg, dfit, per_sample[B,6], mask_losses)."""
    c = FLAGS
    T = lambda v, dt=torch.float32: (v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v))).to(device=dev, dtype=dt).contiguous()
    B = F_output.shape[0]
    NH, NS = NUM_HEADING_BIN, NUM_SIZE_CLUSTER
    E = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    if mean_size is None:
        mean_size, orient_anchors = _consts(dev)
    if F_reg is None:
        F_reg = tf_util.parse_box_output(F_output, stage1_center, mean_size, orient_anchors, want_reg=True)['reg']
    if logits is not None:
        N = logits.shape[1]
        y_seg = T(feed['labels'], torch.int32)
        mask_losses = E(B)
        call('t3d_seg_ce', ptr(logits), ptr(y_seg), B, N, ptr(mask_losses), stream())
    cls_sum, cls_cnt = E(NUM_CLASS, 3), E(NUM_CLASS)
    stat_dims = F_reg[1] if reg_in is None else reg_in[:, 3:6].contiguous()
    call('t3d_class_dims_stats', ptr(stat_dims), ptr(one_hot), B, NUM_CLASS, ptr(cls_sum), ptr(cls_cnt), stream())
    is2d = feed['is_data_2D']
    is2d_host = is2d.cpu().numpy() if torch.is_tensor(is2d) else np.asarray(is2d)
    inv_n3d = 1.0 / (float((1 - is2d_host.astype(np.int64)).sum()) + 1e-3)
    lab = dict(y_center=T(feed['centers']), y_orient_cls=T(feed['y_orient_cls'], torch.int32), y_orient_reg=T(feed['y_orient_reg']),
               y_dims_cls=T(feed['y_dims_cls'], torch.int32), y_dims_reg=T(feed['y_dims_reg']), Rtilt=T(feed['Rtilt']),
               K=T(feed['K']), rot_frust=T(feed['rot_frust']).reshape(B), box2D=T(feed['box2D']), img_dim=T(feed['img_dim']),
               is_data_2D=T(is2d, torch.int32))
    W = 3 + 2 * NH + 4 * NS
    dF, ds1, g_reg, dfit, per_sample, total = E(B, W), E(B, 3), E(B, 7), E(B, 2), E(B, 6), E(8)
    tb = c.WEAK_TRAIN_BOX_W_REPROJECTION
    a = t3d_semi_loss_args()
    keep = dict(out=F_output, stage1_center=stage1_center, mask_losses=mask_losses, one_hot=one_hot, fit_logits=fit_logits,
                mean_size=mean_size, cls_sum=cls_sum, cls_cnt=cls_cnt, dF=dF, ds1=ds1, g_reg=g_reg, dfit=dfit,
                per_sample=per_sample, total=total, reg_in=reg_in, **lab)
    for k, v in keep.items():
        setattr(a, k, ptr(v))
    a.B, a.NH, a.NS, a.NC = B, NH, NS, NUM_CLASS
    a.icv_train_mask = icv_train_mask(c) if icv_mask is None else icv_mask
    a.w_ce, a.box_mult = float(c.STRONG_WEIGHT_CROSS_ENTROPY), float(c.STRONG_BOX_MULTIPLER)
    a.w_center, a.w_ocls, a.w_dcls = float(c.STRONG_WEIGHT_CENTER), float(c.STRONG_WEIGHT_ORIENT_CLS), float(c.STRONG_WEIGHT_DIMS_CLS)
    a.w_oreg, a.w_dreg = float(c.STRONG_WEIGHT_ORIENT_REG), float(c.STRONG_WEIGHT_DIMS_REG)
    a.w_tnet, a.w_corner = float(c.STRONG_WEIGHT_TNET_CENTER), float(c.STRONG_WEIGHT_CORNER)
    a.weak_mult, a.w_icv = float(c.SEMI_MULTIPLIER_FOR_WEAK_LOSS), float(c.WEAK_WEIGHT_INTRACLASSVAR)
    a.w_reproj = float(c.WEAK_WEIGHT_REPROJECTION)
    a.w_fit = float(c.SEMI_WEIGHT_BOXPC_FIT_LOSS) if fit_logits is not None else 0.0
    a.reproj_only_2d, a.fit_only_2d = int(bool(c.WEAK_REPROJECTION_ONLY_ON_2D_CLS)), int(bool(c.SEMI_BOXPC_FIT_ONLY_ON_2D_CLS))
    a.use_softmax_proj, a.softmax_scale = int(bool(c.WEAK_REPROJECTION_USE_SOFTMAX_PROJ)), float(c.WEAK_REPROJECTION_SOFTMAX_SCALE)
    a.dilate = float(c.WEAK_REPROJECTION_DILATE_FACTOR)
    a.clip_lower_b, a.clip_pred_box = int(bool(c.WEAK_REPROJECTION_CLIP_LOWERB_LOSS)), int(bool(c.WEAK_REPROJECTION_CLIP_PRED_BOX))
    a.reproj_mse, a.icv_mse = int(c.WEAK_REPROJECTION_LOSS_TYPE == 'mse'), int(c.WEAK_DIMS_LOSS_TYPE == 'mse')
    a.train_box_mask = (1 if tb[0] else 0) | (2 if tb[1] else 0) | (4 if tb[2] else 0)
    a.inv_n3d = inv_n3d
    if model_a:
        a.inv_n3d, a.reproj_only_2d, a.w_icv, a.w_fit = 1.0 / B, 1, 0.0, 0.0
    call('t3d_semi_loss', ctypes.byref(a), stream())
    iv_out = None
    if c.WEAK_WEIGHT_INACTIVE_VOLUME != 0 and not model_a:           # semisup_v1_sunrgbd.py:348-360: folded into total / weak_loss / g_reg
        assert len(c.WEAK_INACTIVE_VOL_LOSS_MARGINS) == NUM_CLASS
        margins = T(np.asarray(c.WEAK_INACTIVE_VOL_LOSS_MARGINS, dtype=np.float32))
        iv_out = E(1)
        call('t3d_inactive_volume_loss', ptr(stat_dims), ptr(one_hot), ptr(margins), B, NUM_CLASS, inactive_vol_train_mask(c),
             float(c.WEAK_WEIGHT_INACTIVE_VOLUME), float(c.SEMI_MULTIPLIER_FOR_WEAK_LOSS), ptr(iv_out), ptr(total), ptr(g_reg), stream())
    res = dict(inactive_vol=iv_out, total=total, dF=dF, ds1=ds1, g_reg=g_reg, dfit=dfit, per_sample=per_sample, mask_losses=mask_losses, F_reg=F_reg,
               _mean_size=mean_size, _F_output=F_output)
    if finish:
        finish_box_reg(res)
    return res


def finish_box_reg(res):
    """dF += chain of g_reg through the anchor->reg conversion; ds1 += g_reg[:, 0:3]."""
    F_output = res['_F_output']
    call('t3d_box_reg_backward', ptr(F_output), ptr(res['g_reg']), ptr(res['_mean_size']), F_output.shape[0], NUM_HEADING_BIN,
         NUM_SIZE_CLUSTER, ptr(res['dF']), ptr(res['ds1']), stream())
