"""Oracle of the semi-supervised ("adv") training step: restatement of
sunrgbd_detection/train_semisup_adv.py:135-153 (schedules), :267-425 (graph: model F in training mode ->
frozen BoxPC branch on F_pred_box_reg -> refine loop / F2_* end points -> get_semi_loss -> Adam over
class_dependent + class_agnostic/tnet + class_agnostic/box*) with PyTorch autograd for the backward and
TensorFlow's Adam rule.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline may import this.

What the reference builds but the step never evaluates (TF prunes it from the fetches) is not restated: the
real/fake "D" branches and D_loss (:326-354; never minimised, in no summary), get_iou_summary (:414-416,
metrics-only tf.py_func around the missing box_util.box3d_iou).  The "fake" BoxPC branch on the un-refined box
(:339-343) IS restated: its fit probability feeds the loss unless SEMI_BOXPC_MIN_FIT_LOSS_AFT_REFINE (:355, :390-391;
the two differ as soon as SEMI_REFINE_USING_BOXPC_DELTA_NUM > 1 -- found by running the reference's own train() graph,
tests/golden/reference_cases.py).
"""
import numpy as np
import torch

from . import semisup_v1_sunrgbd as MODEL
from .tf_layers import VarStore
from .train_boxpc import get_learning_rate, get_bn_decay, adam_step_tf   # identical schedules (:135-153)

ALL_CLASSES = ['bed', 'table', 'sofa', 'chair', 'toilet', 'desk', 'dresser', 'night_stand', 'bookshelf', 'bathtub']
SUNRGBD_SEMI_TEST_CLS = ['table', 'sofa', 'dresser', 'night_stand', 'bookshelf']      # models/config.py:193-194


def train_var_prefixes(FLAGS):
    """train_semisup_adv.py:415-419: get_scope_vars is a prefix match (tf.get_collection(scope=...))."""
    p = ['class_dependent']
    if getattr(FLAGS, 'SEMI_TRAIN_BOX_TRAIN_CLASS_AG_TNET', False):
        p.append('class_agnostic/tnet')
    if getattr(FLAGS, 'SEMI_TRAIN_BOX_TRAIN_CLASS_AG_BOX', False):
        p.append('class_agnostic/box')
    return tuple(p)


def class_lists(FLAGS):
    """train_semisup_adv.py:320-323."""
    test_cls = getattr(FLAGS, 'TEST_CLS', SUNRGBD_SEMI_TEST_CLS)
    only2d = [c in test_cls for c in ALL_CLASSES]
    icv = only2d if getattr(FLAGS, 'SEMI_INTRACLSDIMS_ONLY_ON_2D_CLS', True) else [True] * len(ALL_CLASSES)
    iv = only2d if getattr(FLAGS, 'WEAK_INACTIVE_VOL_ONLY_ON_2D_CLS', True) else [True] * len(ALL_CLASSES)
    return icv, iv


def loss_and_grads(variables, FLAGS, feed, dropout_masks, global_step=0, dtype=torch.float32, extra_grads=()):
    """One forward + backward of the step on one batch.
    feed: dict keyed like semisup_v1_sunrgbd.placeholder_inputs (synth.make_batch);
    dropout_masks: {'class_agnostic/inst_seg/dp1': (B,N,128), 'class_dependent/box_refine/dp0': (B,512), '.../dp1': (B,256)}.
    Returns (loss, {trainable var: grad or None}, VarStore with the updated moving statistics, end_points)."""
    prefixes = train_var_prefixes(FLAGS)
    vs = VarStore(variables, dtype=dtype, requires_grad=False)
    for k, t in vs.vars.items():
        if k.startswith(prefixes) and not k.endswith(('moving_mean', 'moving_variance')):
            t.requires_grad_(True)
    for k, v in dropout_masks.items():
        vs.dropout_masks[k] = torch.as_tensor(np.asarray(v)).to(dtype)
    T = lambda v, dt=dtype: torch.as_tensor(np.asarray(v)).to(dt)
    pc, one_hot = T(feed['pc']), T(feed['one_hot'])
    B = pc.shape[0]
    saved_refine = FLAGS.refine
    FLAGS.refine = FLAGS.SEMI_REFINE_USING_BOXPC_DELTA_NUM
    try:
        # tf_normalize_2D_bboxes + get_semi_model + BoxPC loop + F2_* (train_semisup_adv.py:314-411), G in training mode,
        # BoxPC in eval mode (is_training_D = False unless SEMI_TRAIN_BOXPC_MODEL)
        assert not FLAGS.SEMI_TRAIN_BOXPC_MODEL
        bn_decay = get_bn_decay(global_step, B)
        logits, ep = _run_graph_training(vs, FLAGS, pc, one_hot, T(feed['box2D']), T(feed['img_dim']), bn_decay)
    finally:
        FLAGS.refine = saved_refine
    icv_cls, iv_cls = class_lists(FLAGS)
    ep['intraclsdims_train_classes'] = icv_cls
    ep['inactive_vol_train_classes'] = iv_cls
    I = torch.int64
    labels = (T(feed['labels'], I), T(feed['centers']), T(feed['y_orient_cls'], I), T(feed['y_orient_reg']),
              T(feed['y_dims_cls'], I), T(feed['y_dims_reg']), None, None, T(feed['Rtilt']), T(feed['K']),
              T(feed['rot_frust']), T(feed['box2D']), T(feed['img_dim']), T(feed['is_data_2D'], I))
    pred = (logits, ep['_W_pred_box'], ep['_F_pred_box'])
    loss = MODEL.get_semi_loss(pred, labels, ep, c=FLAGS)
    names = [k for k, v in vs.vars.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [vs.vars[k] for k in names] + [ep[k] for k in extra_grads], allow_unused=True)
    for k, g in zip(extra_grads, grads[len(names):]):
        ep['d_' + k] = g            # gradients w.r.t. intermediate end points (test diagnostics)
    return loss.detach(), {k: g for k, g in zip(names, grads)}, vs, ep


def _run_graph_training(vs, FLAGS, pc, one_hot, box2D, img_dim, bn_decay):
    """test_semisup.run_graph with is_training=True and the bn_decay schedule threaded through."""
    from . import tf_util, boxpc_sunrgbd
    norm_box2D = tf_util.tf_normalize_2D_bboxes(box2D, img_dim)
    pred, end_points = MODEL.get_semi_model(pc, None, None, one_hot, True, FLAGS.use_one_hot, vs,
                                            norm_box2D=norm_box2D, bn_decay=bn_decay, c=FLAGS)
    logits = pred[0]
    end_points['_W_pred_box'], end_points['_F_pred_box'] = pred[1], pred[2]
    curr_box = end_points['F_pred_box_reg']
    tot_c = torch.zeros_like(curr_box[0])
    tot_s = torch.zeros_like(curr_box[1])
    tot_a = torch.zeros_like(curr_box[2])
    # the "fake" branch on the un-refined box (:339-343, :355): its fit probability is the one the loss sees unless
    # SEMI_BOXPC_MIN_FIT_LOSS_AFT_REFINE re-reads it after the loop (:390-391); with one refinement step both are the same tensor
    with vs.variable_scope('D_boxpc_branch'):
        _, ep = boxpc_sunrgbd.get_model((curr_box, pc), False, one_hot, vs,
                                        use_one_hot_vec=getattr(FLAGS, 'use_one_hot_boxpc', False), c=FLAGS)
    fake_fit_logits = ep['boxpc_fit_logits']
    for _ in range(int(FLAGS.SEMI_REFINE_USING_BOXPC_DELTA_NUM)):
        with vs.variable_scope('D_boxpc_branch'):
            _, ep = boxpc_sunrgbd.get_model((curr_box, pc), False, one_hot, vs,
                                            use_one_hot_vec=getattr(FLAGS, 'use_one_hot_boxpc', False), c=FLAGS)
        w = (1 - ep['logits_for_weigh']) if FLAGS.SEMI_WEIGH_BOXPC_DELTA_DURING_TEST else torch.ones_like(ep['logits_for_weigh'])
        dc, da, ds = ep['boxpc_delta_center'] * w.unsqueeze(1), ep['boxpc_delta_angle'] * w, ep['boxpc_delta_size'] * w.unsqueeze(1)
        curr_box = (curr_box[0] - dc, curr_box[1] - ds, curr_box[2] - da)
        tot_c, tot_s, tot_a = tot_c + dc, tot_s + ds, tot_a + da
    if FLAGS.SEMI_BOXPC_MIN_FIT_LOSS_AFT_REFINE:
        fake_fit_logits = ep['boxpc_fit_logits']
    end_points['boxpc_fit_prob'] = torch.softmax(fake_fit_logits, dim=1)[:, 1]
    end_points['boxpc_fit_logits'] = fake_fit_logits
    end_points['boxpc_delta_center'] = ep['boxpc_delta_center']
    end_points['boxpc_delta_size'] = ep['boxpc_delta_size']
    end_points['boxpc_delta_angle'] = ep['boxpc_delta_angle']
    end_points['F2_center'] = end_points['F_center'] - tot_c
    end_points['F2_heading_residuals'] = end_points['F_heading_residuals'] - tot_a.unsqueeze(1).repeat(1, 12)
    end_points['F2_size_residuals'] = end_points['F_size_residuals'] - tot_s.unsqueeze(1).repeat(1, 10, 1)
    end_points['logits'] = logits
    return logits, end_points
