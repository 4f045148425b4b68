"""Oracle of the BoxPC training step: restatement of sunrgbd_detection/train_boxpc.py:134-152 (schedules),
:229-256 (graph: convert_raw_y_box_to_reg_format -> get_model(is_training) -> get_loss -> Adam) with PyTorch
autograd for the backward and TensorFlow's Adam update rule."""
import numpy as np
import torch

from . import boxpc_sunrgbd
from .tf_layers import VarStore


def get_learning_rate(batch, batch_size, base_learning_rate=0.001, decay_step=800000, decay_rate=0.5):
    """train_boxpc.py:134-142 (staircase; the clip is a no-op typo in the reference)."""
    return base_learning_rate * decay_rate ** ((batch * batch_size) // decay_step)


def get_bn_decay(batch, batch_size, decay_step=800000):
    """train_boxpc.py:144-152."""
    return min(0.99, 1 - 0.5 * 0.5 ** ((batch * batch_size) // int(decay_step)))


def loss_and_grads(variables, FLAGS, feed, dropout_masks, global_step=0, dtype=torch.float32):
    """Returns (loss, {var: grad}, updated VarStore (moving stats), end_points)."""
    vs = VarStore(variables, dtype=dtype, requires_grad=True)
    for k, v in dropout_masks.items():
        vs.dropout_masks['box_pc_mask_model/' + k] = torch.as_tensor(np.asarray(v)).to(dtype)
    T = lambda v, dt=dtype: torch.as_tensor(np.asarray(v)).to(dt)
    one_hot = T(feed['one_hot'])
    x_box = (T(feed['x_center']), T(feed['x_orient_cls'], torch.int64), T(feed['x_orient_reg']),
             T(feed['x_dims_cls'], torch.int64), T(feed['x_dims_reg']))
    box_reg = boxpc_sunrgbd.convert_raw_y_box_to_reg_format(x_box, one_hot)
    B = one_hot.shape[0]
    bn_decay = get_bn_decay(global_step, B)
    pred, ep = boxpc_sunrgbd.get_model((box_reg, T(feed['pc'])), True, one_hot, vs, use_one_hot_vec=False,
                                       bn_decay=bn_decay, c=FLAGS)
    labels = (T(feed['y_box_iou']), (T(feed['y_center_delta']), T(feed['y_dims_delta']), T(feed['y_orient_delta'])))
    loss = boxpc_sunrgbd.get_loss(pred, labels, ep, c=FLAGS)
    names = [k for k, v in vs.vars.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [vs.vars[k] for k in names])
    return loss.detach(), {k: g for k, g in zip(names, grads)}, vs, ep


def adam_step_tf(param, grad, m, v, lr, t, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps)."""
    m = beta1 * m + (1 - beta1) * grad
    v = beta2 * v + (1 - beta2) * grad * grad
    lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    return param - lr_t * m / (torch.sqrt(v) + eps), m, v
