"""Oracle of the model-A semi-supervised training step: restatement of sunrgbd_detection/train_semisup.py:199-262
(train(): placeholders -> tf_normalize_2D_bboxes -> get_semi_model (SEMI_MODEL 'A', is_training) -> get_semi_loss ->
optimizer.minimize(semi_loss, global_step=batch) over ALL trainable variables) with PyTorch autograd for the backward.
Test infrastructure: only tests/ may import it."""
import numpy as np
import torch

from . import semisup_v1_sunrgbd as MODEL
from . import tf_util
from .tf_layers import VarStore
from .train_boxpc import get_bn_decay, get_learning_rate, adam_step_tf  # noqa: F401 (same schedules, train_semisup.py:118-136)


def loss_and_grads(variables, FLAGS, feed, dropout_masks, global_step=0, dtype=torch.float32):
    """feed: dict keyed like semisup_v1_sunrgbd.placeholder_inputs (synth.make_batch); dropout_masks: {'inst_seg/dp1': (B,N,128)}.
    Returns (loss, {var: grad or None}, VarStore with the updated moving statistics, end_points)."""
    assert FLAGS.SEMI_MODEL == 'A'
    vs = VarStore(variables, dtype=dtype, requires_grad=True)
    for k, v in dropout_masks.items():
        vs.dropout_masks[k] = torch.as_tensor(np.asarray(v)).to(dtype)
    T = lambda v, dt=dtype: torch.as_tensor(np.asarray(v)).to(dt)
    pc, one_hot = T(feed['pc']), T(feed['one_hot'])
    B = pc.shape[0]
    bn_decay = get_bn_decay(global_step, B)
    norm_box2D = tf_util.tf_normalize_2D_bboxes(T(feed['box2D']), T(feed['img_dim']))
    pred, ep = MODEL.get_semi_model(pc, None, None, one_hot, True, FLAGS.use_one_hot, vs, norm_box2D=norm_box2D,
                                    bn_decay=bn_decay, c=FLAGS)
    I = torch.int64
    labels = (T(feed['labels'], I), T(feed['centers']), T(feed['y_orient_cls'], I), T(feed['y_orient_reg']),
              T(feed['y_dims_cls'], I), T(feed['y_dims_reg']), None, None, T(feed['Rtilt']), T(feed['K']),
              T(feed['rot_frust']), T(feed['box2D']), T(feed['img_dim']), T(feed['is_data_2D'], I))
    loss = MODEL.get_semi_loss(pred, labels, ep, c=FLAGS)
    names = [k for k, v in vs.vars.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [vs.vars[k] for k in names], allow_unused=True)
    ep['logits'] = pred[0]
    return loss.detach(), {k: g for k, g in zip(names, grads)}, vs, ep
