"""Mirror of the BoxPC-Fit training graph and step of sunrgbd/sunrgbd_detection/train_boxpc.py
(get_learning_rate :134-142, get_bn_decay :144-152, train :219-300, the sess.run of train_one_epoch :346-366)
on the B200: forward in training mode (batch-statistics BN, dropout with explicit keep masks), backward,
one NCCL all-reduce of the flat fp32 gradient arena when world_size > 1, TF-style Adam -- all in
libt3d_b200.so kernels (fp32 CUDA-core path in this round).

BN statistics and dropout are per replica (the reference is single-device; parity target = one replica at its
local batch).
"""
import ctypes

import numpy as np
import torch

from . import runtime as rt
from . import boxpc_sunrgbd
from ._lib import ptr, stream, call, t3d_boxpc_loss_args
from .constants import BN_EPS
from .weights import net_table
from .train_layers import TrainLayer, ACT_RELU, ACT_NONE, maxpool

BN_INIT_DECAY = 0.5
BN_DECAY_DECAY_RATE = 0.5
BN_DECAY_CLIP = 0.99


def get_learning_rate(batch, batch_size, base_learning_rate=0.001, decay_step=800000, decay_rate=0.5):
    """train_boxpc.py:134-142: staircase exponential decay (the clip line there is a no-op typo)."""
    return base_learning_rate * decay_rate ** ((batch * batch_size) // decay_step)


def get_bn_decay(batch, batch_size, decay_step=800000):
    """train_boxpc.py:144-152."""
    bn_momentum = BN_INIT_DECAY * BN_DECAY_DECAY_RATE ** ((batch * batch_size) // int(decay_step))
    return min(BN_DECAY_CLIP, 1 - bn_momentum)


class BoxPCTrainGraph(object):
    """What train_boxpc.train() builds: placeholders -> convert_raw_y_box_to_reg_format -> boxpc get_model
    (is_training) -> get_loss -> Adam.minimize over ALL variables (train_boxpc.py:229-256)."""

    def __init__(self, variables, FLAGS, batch_size, num_point, num_channels=6, device='cuda', base_learning_rate=0.001,
                 decay_step=800000, decay_rate=0.5, scope='box_pc_mask_model', process_group=None):
        self.FLAGS, self.B, self.Npt, self.C = FLAGS, batch_size, num_point, num_channels
        self.device = torch.device(device)
        self.base_lr, self.decay_step, self.decay_rate = base_learning_rate, decay_step, decay_rate
        self.scope = scope
        self.pg = process_group
        self.use_one_hot = bool(getattr(FLAGS, 'use_one_hot_boxpc', False))
        table = net_table('box_pc_mask_model', num_channels, one_hot=self.use_one_hot)
        # flat arenas (parameters, gradients, Adam moments); moving statistics are not trainable
        names, sizes = [], []
        for lname, kind, kw, cin, cout, bn in table:
            for suf in ('weights', 'biases') + (('bn/gamma', 'bn/beta') if bn else ()):
                full = '%s/%s/%s' % (scope, lname, suf)
                names.append(full)
                sizes.append(int(np.prod(variables[full].shape)))
        total = sum(sizes)
        self.flat_param = torch.empty(total, dtype=torch.float32, device=self.device)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.adam_m = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.adam_v = torch.zeros(total, dtype=torch.float32, device=self.device)
        self.param, self.grad, self.moving = {}, {}, {}
        off = 0
        for n, sz in zip(names, sizes):
            self.param[n[len(scope) + 1:]] = self.flat_param[off:off + sz]
            self.grad[n[len(scope) + 1:]] = self.flat_grad[off:off + sz]
            self.flat_param[off:off + sz].copy_(torch.as_tensor(np.asarray(variables[n], dtype=np.float32).reshape(-1)))
            off += sz
        for lname, kind, kw, cin, cout, bn in table:
            if bn:
                for suf in ('bn/moving_mean', 'bn/moving_variance'):
                    self.moving['%s/%s' % (lname, suf)] = torch.as_tensor(
                        np.asarray(variables['%s/%s/%s' % (scope, lname, suf)], dtype=np.float32)).to(self.device).contiguous()
        self.layers = []
        for i, (lname, kind, kw, cin, cout, bn) in enumerate(table):
            self.layers.append(TrainLayer(lname, kw * cin if kind == 'conv' else cin, cout, bn, ACT_RELU if bn else ACT_NONE,
                                          self.param, self.moving, self.grad))
        self.global_step = 0
        self._store = rt.VariableStore({}, self.device)        # constants only (anchors)

    # ------------------------------------------------------------------------------------------
    def variables(self):
        """Current values keyed by TF variable name (what tf.train.Saver would write)."""
        out = {}
        for k, v in self.param.items():
            out['%s/%s' % (self.scope, k)] = v.detach().cpu().numpy().copy()
        for k, v in self.moving.items():
            out['%s/%s' % (self.scope, k)] = v.detach().cpu().numpy().copy()
        return out

    def forward_backward(self, feed, dropout_masks):
        """One forward + backward in training mode. feed: dict keyed like boxpc_sunrgbd.placeholder_inputs;
        dropout_masks: {'dp1': (B,512) keep mask, 'dp2': (B,256)}. Leaves gradients in self.grad."""
        dev = self.device
        T = lambda v, dt=torch.float32: (v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v))).to(device=dev, dtype=dt).contiguous()
        B, N, C = self.B, self.Npt, self.C
        pc = T(feed['pc'])
        one_hot = T(feed['one_hot'])
        rt.set_default_store(self._store)
        x_box = (T(feed['x_center']), T(feed['x_orient_cls'], torch.int32), T(feed['x_orient_reg']),
                 T(feed['x_dims_cls'], torch.int32), T(feed['x_dims_reg']))
        box_reg = boxpc_sunrgbd.convert_raw_y_box_to_reg_format(x_box, one_hot)
        bn_decay = get_bn_decay(self.global_step, B, self.decay_step)
        from . import tf_util
        rep = tf_util.tf_get_box_pc_representation(box_reg, pc).reshape(B * N, C + 6)
        L = self.layers
        x = rep
        for l in L[:4]:
            x = l.forward(x, bn_decay, lazy=True)       # lazy BN: statistics from the GEMM epilogue, BN map applied by the consumer
        pooled, self.arg = maxpool(x, B, N, 512)
        feat = torch.cat([pooled, one_hot], dim=1).contiguous() if self.use_one_hot else pooled
        h1 = L[4].forward(feat, bn_decay)
        m1 = T(dropout_masks['dp1'])
        d1 = torch.empty_like(h1)
        call('t3d_scale_mask', ptr(h1), ptr(m1), 1.0 / 0.7, ptr(d1), h1.numel(), stream())
        h2 = L[5].forward(d1, bn_decay)
        m2 = T(dropout_masks['dp2'])
        d2 = torch.empty_like(h2)
        call('t3d_scale_mask', ptr(h2), ptr(m2), 1.0 / 0.7, ptr(d2), h2.numel(), stream())
        out9 = L[6].forward(d2, bn_decay)
        # loss + d loss / d out9
        c = self.FLAGS
        total = torch.empty(1, device=dev)
        g9 = torch.empty((B, 9), device=dev)
        cls_l = torch.empty(B, device=dev)
        del_l = torch.empty(B, device=dev)
        # keep the label tensors referenced until the kernel is enqueued (their memory must not be recycled)
        y_iou, y_dc, y_ds, y_da = T(feed['y_box_iou']), T(feed['y_center_delta']), T(feed['y_dims_delta']), T(feed['y_orient_delta'])
        a = t3d_boxpc_loss_args(ptr(out9), ptr(y_iou), ptr(y_dc), ptr(y_ds), ptr(y_da), B, float(c.BOXPC_FIT_BOUNDS[0]), float(c.BOXPC_WEIGHT_CLS),
                                float(c.BOXPC_WEIGHT_DELTA), float(c.BOXPC_WEIGHT_DELTA_CENTER_PERCENT),
                                float(c.BOXPC_WEIGHT_DELTA_SIZE_PERCENT), float(c.BOXPC_WEIGHT_DELTA_ANGLE_PERCENT),
                                1 if c.BOXPC_DELTA_LOSS_TYPE == 'huber' else 0, ptr(cls_l), ptr(del_l), ptr(total), ptr(g9))
        if c.BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF or c.BOXPC_WEIGH_DELTA_LOSS_BY_CLS_GT or c.BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF:
            raise NotImplementedError('BOXPC_WEIGH_DELTA_* options are not on the recipe path (scripts/train_semisup_bed.sh)')
        call('t3d_boxpc_loss', ctypes.byref(a), stream())
        # backward
        g = L[6].backward(g9)
        gd = torch.empty_like(g)
        call('t3d_scale_mask', ptr(g), ptr(m2), 1.0 / 0.7, ptr(gd), g.numel(), stream())
        g = L[5].backward(gd)
        gd = torch.empty_like(g)
        call('t3d_scale_mask', ptr(g), ptr(m1), 1.0 / 0.7, ptr(gd), g.numel(), stream())
        g = L[4].backward(gd)
        if self.use_one_hot:
            g = g[:, :512].contiguous()
        g = L[3].backward_pooled(g.contiguous(), self.arg, B, N)     # max-pool + BN backward without the dense pooled gradient
        g = L[2].backward(g)
        g = L[1].backward(g)
        L[0].backward(g, need_dx=False)
        fit_prob = torch.softmax(out9[:, 7:9], dim=1)[:, 1]
        return {'loss': total, 'boxpc_cls_losses': cls_l, 'boxpc_delta_losses': del_l, 'output': out9,
                'pred_boxpc_fit': (fit_prob > 0.5).to(torch.int32), 'boxpc_delta_center': out9[:, 0:3],
                'boxpc_delta_size': out9[:, 3:6], 'boxpc_delta_angle': out9[:, 6]}

    def apply_gradients(self):
        """Adam over every variable (optimizer.minimize(loss, global_step=batch), train_boxpc.py:249-256)."""
        from .dist_util import allreduce_flat
        world = allreduce_flat(self.flat_grad, self.pg)                              # one NCCL all-reduce of the flat arena
        lr = get_learning_rate(self.global_step, self.B, self.base_lr, self.decay_step, self.decay_rate)
        t = self.global_step + 1
        b1, b2 = 0.9, 0.999
        lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        call('t3d_adam', ptr(self.flat_param), ptr(self.flat_grad), ptr(self.adam_m), ptr(self.adam_v), self.flat_param.numel(),
             float(lr_t), b1, b2, 1e-8, 1.0 / world, stream())
        self.global_step += 1

    def step(self, feed, dropout_masks):
        """sess.run([..., loss, train_op, pred_boxpc_fit, deltas]) of train_one_epoch (train_boxpc.py:360-366)."""
        out = self.forward_backward(feed, dropout_masks)
        self.apply_gradients()
        out['step'] = self.global_step
        return out
