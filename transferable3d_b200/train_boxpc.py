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


def _scale_mask(x, mask, scale):
    out = torch.empty_like(x)
    call('t3d_scale_mask', ptr(x), ptr(mask), float(scale), ptr(out), x.numel(), stream())
    return out


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
        self.rep = str(getattr(FLAGS, 'BOX_PC_MASK_REPRESENTATION', 'A'))
        if self.rep not in ('A', 'B'):
            raise Exception('Box pc mask representation not implemented: %s' % self.rep)
        table = net_table('box_pc_mask_model' if self.rep == 'A' else 'box_pc_mask_model_B', num_channels, one_hot=self.use_one_hot)
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
        dropout_masks: {'dp1': (B,512) keep mask, 'dp2': (B,256)} (representation B: {'dp2': (B,512), 'dp3': (B,256)}).
        Leaves gradients in self.grad."""
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
        L = self.layers
        drop = lambda h, m: _scale_mask(h, m, 1.0 / 0.7)
        if self.rep == 'A':
            # semisup_models.py:326-398: conv stack on [pc, 6 plane distances], max, fc1 - dp1 - fc2 - dp2 - fc3
            conv, head, dps = L[:4], L[4:], ('dp1', 'dp2')
            x = tf_util.tf_get_box_pc_representation(box_reg, pc).reshape(B * N, C + 6)
        else:
            # semisup_models.py:400-471: box (B,7) through extract_box_feats, conv stack on the raw points, max,
            # [box_feat, point_feat] through fc1 - fc2 - dp2 - fc3 - dp3 - fc4
            boxfc, conv, head, dps = L[:4], L[4:8], L[8:], ('dp2', 'dp3')
            h = torch.cat([box_reg[0], box_reg[1], box_reg[2].reshape(B, 1)], dim=1).contiguous()
            for l in boxfc:
                h = l.forward(h, bn_decay)
            box_feat = h
            x = pc.reshape(B * N, C)
        for l in conv:
            x = l.forward(x, bn_decay, lazy=True)       # lazy BN: statistics from the GEMM epilogue, BN map applied by the consumer
        pooled, self.arg = maxpool(x, B, N, 512)
        parts = ([pooled] if self.rep == 'A' else [box_feat, pooled]) + ([one_hot] if self.use_one_hot else [])
        feat = torch.cat(parts, dim=1).contiguous() if len(parts) > 1 else pooled
        m1, m2 = T(dropout_masks[dps[0]]), T(dropout_masks[dps[1]])
        if self.rep == 'A':
            h1 = head[0].forward(feat, bn_decay)
            h2 = head[1].forward(drop(h1, m1), bn_decay)
            out9 = head[2].forward(drop(h2, m2), bn_decay)
        else:
            h1 = head[0].forward(feat, bn_decay)
            h2 = head[1].forward(h1, bn_decay)
            h3 = head[2].forward(drop(h2, m1), bn_decay)
            out9 = head[3].forward(drop(h3, m2), bn_decay)
        # loss + d loss / d out9 (the RAW network output: the class-confidence weighting of the deltas is inside the kernel)
        c = self.FLAGS
        total = torch.empty(1, device=dev)
        g9 = torch.empty((B, 9), device=dev)
        cls_l = torch.empty(B, device=dev)
        del_l = torch.empty(B, device=dev)
        # keep the label tensors referenced until the kernel is enqueued (their memory must not be recycled)
        y_iou, y_dc, y_ds, y_da = T(feed['y_box_iou']), T(feed['y_center_delta']), T(feed['y_dims_delta']), T(feed['y_orient_delta'])
        assert not (c.BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF and c.BOXPC_WEIGH_DELTA_LOSS_BY_CLS_GT)
        pred_weigh = 1 if c.BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF else 0
        loss_weigh = 1 if c.BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF else (2 if c.BOXPC_WEIGH_DELTA_LOSS_BY_CLS_GT else 0)
        a = t3d_boxpc_loss_args(ptr(out9), ptr(y_iou), ptr(y_dc), ptr(y_ds), ptr(y_da), B, float(c.BOXPC_FIT_BOUNDS[0]), float(c.BOXPC_WEIGHT_CLS),
                                float(c.BOXPC_WEIGHT_DELTA), float(c.BOXPC_WEIGHT_DELTA_CENTER_PERCENT),
                                float(c.BOXPC_WEIGHT_DELTA_SIZE_PERCENT), float(c.BOXPC_WEIGHT_DELTA_ANGLE_PERCENT),
                                1 if c.BOXPC_DELTA_LOSS_TYPE == 'huber' else 0, ptr(cls_l), ptr(del_l), ptr(total), ptr(g9),
                                pred_weigh, loss_weigh, 1 if c.BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA else 0)
        call('t3d_boxpc_loss', ctypes.byref(a), stream())
        # backward
        if self.rep == 'A':
            g = head[2].backward(g9)
            g = head[1].backward(drop(g, m2))
            g = head[0].backward(drop(g, m1))
            g_pool = g[:, :512].contiguous() if self.use_one_hot else g
        else:
            g = head[3].backward(g9)
            g = head[2].backward(drop(g, m2))
            g = head[1].backward(drop(g, m1))
            g = head[0].backward(g)
            g_box, g_pool = g[:, :512].contiguous(), g[:, 512:1024].contiguous()
        g = conv[3].backward_pooled(g_pool.contiguous(), self.arg, B, N)     # max-pool + BN backward without the dense pooled gradient
        g = conv[2].backward(g)
        g = conv[1].backward(g)
        conv[0].backward(g, need_dx=False)
        if self.rep == 'B':
            g = boxfc[3].backward(g_box)
            g = boxfc[2].backward(g)
            g = boxfc[1].backward(g)
            boxfc[0].backward(g, need_dx=False)
        fit_prob = torch.softmax(out9[:, 7:9], dim=1)[:, 1]
        w = (1.0 - fit_prob) if pred_weigh else torch.ones_like(fit_prob)
        return {'loss': total, 'boxpc_cls_losses': cls_l, 'boxpc_delta_losses': del_l, 'output': out9,
                'pred_boxpc_fit': (fit_prob > 0.5).to(torch.int32), 'boxpc_delta_center': out9[:, 0:3] * w[:, None],
                'boxpc_delta_size': out9[:, 3:6] * w[:, None], 'boxpc_delta_angle': out9[:, 6] * w}

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
