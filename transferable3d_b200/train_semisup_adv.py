"""Mirror of the semi-supervised ("adv") training graph and step of
sunrgbd/sunrgbd_detection/train_semisup_adv.py on the B200 (BASELINE cfg5):

  get_learning_rate / get_bn_decay (:135-153), train() graph (:267-425): model F in training mode
  (semisup_v1_sunrgbd.get_semi_model_final :132-230) -> frozen BoxPC branch on F_pred_box_reg (:337-345,364-391)
  -> F2_* end points (:393-411) -> get_semi_loss (:414 -> semisup_v1_sunrgbd.py:323-421) -> Adam over
  class_dependent + class_agnostic/tnet + class_agnostic/box* (:415-422); the sess.run of train_one_epoch (:604-614).

What runs: seg net forward in training mode (batch-statistics BN, dropout, moving-stat updates; no backward -- it is
not in var_list and the mask is a non-differentiable compare); T-Net, box-est convs and the box_refine head forward +
backward in training mode; the box-est FC head forward only (it feeds the W_ IoU summaries of the step, so its BN
moving statistics move, but no loss term reaches it: TF skips variables whose gradient is None); BoxPC in eval mode
forward + input-gradient only; the losses and their gradients in one O(B) kernel (csrc/loss_ops.cuh).
What the reference builds but its step never evaluates is not run: the real/fake D branches + D_loss (dead), and
get_iou_summary (metrics-only tf.py_func around the missing box_util).

Multi-GPU: data parallel, one all-reduce of the flat gradient arena per step; BN statistics, dropout and the
batch-coupled loss terms (intra-class variance group means, the 3D-sample count) are per replica.
All arithmetic is in libt3d_b200.so kernels (fp32 CUDA-core path in this round); torch ops are used only for O(B)
glue (concat / slice / add).
"""
import ctypes

import numpy as np
import torch

from . import runtime as rt
from . import tf_util, losses
from ._lib import ptr, stream, call, t3d_refine_args
from .constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER, NUM_CLASS, MEAN_DIMS_ARR
from .weights import net_table
from .train_boxpc import get_learning_rate, get_bn_decay          # same schedules (train_semisup_adv.py:135-153)
from .train_layers import (ParamArena, TrainLayer, EvalLayer, ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH, maxpool, maxpool_bwd,
                           rowmask_mul, dropout, gemm, dense, pool_rows, gather_rows, scatter_pool_grad)


def train_var_prefixes(FLAGS):
    """train_semisup_adv.py:415-419 (get_scope_vars matches by prefix, so 'class_agnostic/box' covers box_est)."""
    p = ['class_dependent']
    if FLAGS.SEMI_TRAIN_BOX_TRAIN_CLASS_AG_TNET:
        p.append('class_agnostic/tnet')
    if FLAGS.SEMI_TRAIN_BOX_TRAIN_CLASS_AG_BOX:
        p.append('class_agnostic/box')
    return tuple(p)


class SemiAdvTrainGraph(object):
    def __init__(self, variables, FLAGS, batch_size, num_point, num_channels=6, device='cuda', base_learning_rate=0.001,
                 decay_step=800000, decay_rate=0.5, process_group=None):
        c = FLAGS
        if c.SEMI_MODEL != 'F':
            raise Exception('Not implemented SEMI_MODEL: %s' % c.SEMI_MODEL)
        if c.SEMI_TRAIN_BOXPC_MODEL:
            raise NotImplementedError('SEMI_TRAIN_BOXPC_MODEL: the BoxPC branch is frozen in the recipe (is_training_D = False)')
        if int(c.SEMI_REFINE_USING_BOXPC_DELTA_NUM) != 1:
            raise NotImplementedError('SEMI_REFINE_USING_BOXPC_DELTA_NUM != 1')
        if c.USE_NORMALIZED_BOX2D_AS_FEATS:
            raise NotImplementedError('USE_NORMALIZED_BOX2D_AS_FEATS in the training graph')
        self.FLAGS, self.B, self.Npt, self.C = FLAGS, batch_size, num_point, num_channels
        self.device = dev = torch.device(device)
        self.base_lr, self.decay_step, self.decay_rate = base_learning_rate, decay_step, decay_rate
        self.pg = process_group
        self.global_step = 0
        D = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32)).to(dev).contiguous()

        nets = (('class_agnostic/inst_seg', net_table('inst_seg', num_channels)),
                ('class_agnostic/tnet', net_table('tnet')),
                ('class_agnostic/box_est', net_table('box_est')),
                ('class_dependent/box_refine', net_table('box_refine', one_hot=bool(c.use_one_hot))))
        prefixes = train_var_prefixes(c)
        # variables that are in var_list AND are reached by the loss: everything under the prefixes except the box_est FC head
        train_names, frozen = [], {}
        self.moving = {}
        for scope, table in nets:
            for lname, kind, kw, cin, cout, bn in table:
                layer = '%s/%s' % (scope, lname)
                trainable = layer.startswith(prefixes) and not (scope.endswith('box_est') and kind == 'fc')
                for suf in ('weights', 'biases') + (('bn/gamma', 'bn/beta') if bn else ()):
                    n = '%s/%s' % (layer, suf)
                    if trainable:
                        train_names.append(n)
                    else:
                        frozen[n] = D(variables[n]).reshape(-1)
                if bn:
                    for suf in ('bn/moving_mean', 'bn/moving_variance'):
                        self.moving['%s/%s' % (layer, suf)] = D(variables['%s/%s' % (layer, suf)])
        self.arena = ParamArena(variables, train_names, dev)
        self.param = dict(frozen)
        self.param.update(self.arena.param)
        self.grad = self.arena.grad
        self.train_names = train_names

        def layers(scope, table, acts=None):
            out = []
            for i, (lname, kind, kw, cin, cout, bn) in enumerate(table):
                act = (acts[i] if acts is not None else (ACT_RELU if bn else ACT_NONE))
                out.append(TrainLayer('%s/%s' % (scope, lname), kw * cin if kind == 'conv' else cin, cout, bn, act,
                                      self.param, self.moving, self.grad))
            return out
        self.seg = layers(*nets[0])
        self.tnet = layers(*nets[1])
        self.box = layers(*nets[2])
        a0 = ACT_LEAKY if c.SEMI_ADV_LEAKY_RELU else ACT_RELU
        a1 = ACT_TANH if c.SEMI_ADV_TANH_FOR_LAST_LAYER_OF_G else a0
        self.refine = layers(nets[3][0], nets[3][1], acts=[a0, a1, ACT_NONE])
        # frozen BoxPC branch (eval mode)
        self.boxpc_one_hot = bool(getattr(c, 'use_one_hot_boxpc', False))
        bscope = 'D_boxpc_branch/box_pc_mask_model'
        btable = net_table('box_pc_mask_model', num_channels, one_hot=self.boxpc_one_hot)
        self.boxpc = [EvalLayer(variables, '%s/%s' % (bscope, lname), ACT_RELU if bn else ACT_NONE, dev)
                      for lname, kind, kw, cin, cout, bn in btable]
        self.mean_size = D(MEAN_DIMS_ARR)
        self.orient_anchors = D(np.arange(0, 2 * np.pi, 2 * np.pi / NUM_HEADING_BIN))
        self.icv_mask = losses.icv_train_mask(c)

    # ------------------------------------------------------------------------------------------
    def variables(self):
        """Current values keyed by TF variable name (what tf.train.Saver would write)."""
        out = {k: v.detach().cpu().numpy().copy() for k, v in self.param.items()}
        out.update({k: v.detach().cpu().numpy().copy() for k, v in self.moving.items()})
        return out

    def forward_backward(self, feed, dropout_masks):
        """feed: dict keyed like semisup_v1_sunrgbd.placeholder_inputs (pc, one_hot, labels, centers, y_orient_cls,
        y_orient_reg, y_dims_cls, y_dims_reg, Rtilt, K, rot_frust, box2D, img_dim, is_data_2D);
        dropout_masks: keep masks {'class_agnostic/inst_seg/dp1': (B,N,128), 'class_dependent/box_refine/dp0': (B,512),
        'class_dependent/box_refine/dp1': (B,256)}.  Leaves the gradients in self.grad; returns loss + end points."""
        c, dev = self.FLAGS, self.device
        T = lambda v, dt=torch.float32: (v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v))).to(device=dev, dtype=dt).contiguous()
        B, N, C = self.B, self.Npt, self.C
        NH, NS = NUM_HEADING_BIN, NUM_SIZE_CLUSTER
        pc, one_hot = T(feed['pc']), T(feed['one_hot'])
        assert pc.shape == (B, N, C)
        bn_decay = get_bn_decay(self.global_step, B, self.decay_step)
        ep = {}

        # ---- inst_seg, training mode, forward only (semisup_models.py:69-139); conv6's global half is folded per frustum
        S = self.seg
        x = pc.reshape(B * N, C)
        x = S[0].forward(x, bn_decay, keep=False, lazy=True)
        x = S[1].forward(x, bn_decay, keep=False, lazy=True)
        pf_lazy = S[2].forward(x, bn_decay, keep=False, lazy=True)
        x = S[3].forward(pf_lazy, bn_decay, keep=False, lazy=True)
        gfeat = S[4].forward_pooled(x, bn_decay, B, N)                  # conv5 + BN + ReLU + max-pool: the B*N x 1024 output never exists
        del x
        point_feat = dense(pf_lazy)                                      # [B*N, 64]: the operand of the folded conv6
        W6 = S[5].W()                                                    # [64 + 1024, 512]
        gb = gemm(gfeat, 1024, 1, W6[64:], 512, 1, B, 512, 1024, bias=S[5].p('biases'))
        y6 = torch.empty((B * N, 512), device=dev)
        call('t3d_linear_f32', ptr(point_feat), 64, ptr(W6), 512, None, ptr(gb), N, ptr(y6), 512, B * N, 64, 512, 0, None, None,
             stream())
        x = S[5].forward(None, bn_decay, y=y6, keep=False, lazy=True)
        del y6
        x = S[6].forward(x, bn_decay, keep=False, lazy=True)
        x = S[7].forward(x, bn_decay, keep=False, lazy=True)
        x = S[8].forward(x, bn_decay, keep=False)
        x = dropout(x, T(dropout_masks['class_agnostic/inst_seg/dp1']).reshape(B * N, 128), 0.5)
        logits = S[9].forward(x, bn_decay, keep=False).reshape(B, N, 2)
        del x
        ep['logits'] = logits

        # ---- mask, centroid (semisup_models.py:145-162)
        mask, count, mean, xyz1, _ = rt.mask_centroid(logits, pc, want_mask=True, want_xyz_stage1=True, want_idx=False)
        rowmask = mask.reshape(B * N)

        # ---- T-Net (semisup_models.py:164-202)
        Tn = self.tnet
        x = xyz1.reshape(B * N, 3)
        for l in Tn[:3]:
            x = l.forward(x, bn_decay, lazy=True)
        t_pool, t_arg = maxpool(x, B, N, 256, rowmask)          # max(net * mask) without the product in HBM
        h = Tn[3].forward(t_pool, bn_decay)
        h = Tn[4].forward(h, bn_decay)
        t_out = Tn[5].forward(h, bn_decay)
        stage1_center = (t_out + mean).contiguous()
        ep['stage1_center'] = stage1_center

        # ---- box estimation net (semisup_models.py:204-291); its FC head runs forward only (W_ branch)
        Bx = self.box
        xin = torch.empty((B, N, 3), device=dev)
        call('t3d_prepare_xyz', ptr(pc), B, N, C, ptr(stage1_center), ptr(xin), stream())
        x = xin.reshape(B * N, 3)
        for l in Bx[:4]:
            x = l.forward(x, bn_decay, lazy=True)
        feats_lv1, b_arg = maxpool(x, B, N, 512, rowmask)
        ep['feats_lv1'] = feats_lv1
        h = Bx[4].forward(feats_lv1, bn_decay, keep=False)
        ep['feats_lv2'] = h
        h = Bx[5].forward(h, bn_decay, keep=False)
        ep['feats_lv3'] = h
        ep['box_params'] = Bx[6].forward(h, bn_decay, keep=False)
        wp = tf_util.parse_box_output(ep['box_params'], stage1_center, self.mean_size, self.orient_anchors, want_reg=False)
        for k in ('center', 'heading_scores', 'heading_residuals_normalized', 'heading_residuals', 'size_scores',
                  'size_residuals_normalized', 'size_residuals'):
            ep[k] = wp[k]

        # ---- class-dependent refinement head (semisup_v1_sunrgbd.py:181-222)
        R = self.refine
        keep = float(c.SEMI_ADV_DROPOUTS_FOR_G)
        m0, m1 = T(dropout_masks['class_dependent/box_refine/dp0']), T(dropout_masks['class_dependent/box_refine/dp1'])
        feat = torch.cat([feats_lv1, one_hot], dim=1).contiguous() if c.use_one_hot else feats_lv1
        h = R[0].forward(feat, bn_decay)
        h = dropout(h, m0, keep)
        h = R[1].forward(h, bn_decay)
        h = dropout(h, m1, keep)
        F_output = R[2].forward(h, bn_decay)
        fp = tf_util.parse_box_output(F_output, stage1_center, self.mean_size, self.orient_anchors, want_reg=True)
        for k in ('center', 'heading_scores', 'heading_residuals_normalized', 'heading_residuals', 'size_scores',
                  'size_residuals_normalized', 'size_residuals'):
            ep['F_' + k] = fp[k]
        ep['F_output'] = F_output
        F_reg = fp['reg']
        ep['F_pred_box_reg'] = F_reg

        # ---- frozen BoxPC branch on (F_pred_box_reg, pc), eval mode (train_semisup_adv.py:364-391)
        P = self.boxpc
        rep = tf_util.tf_get_box_pc_representation(F_reg, pc).reshape(B * N, C + 6)
        x = rep
        for l in P[:4]:
            x = l.forward(x)
        p_pool, p_arg = maxpool(x, B, N, 512)
        pf = torch.cat([p_pool, one_hot], dim=1).contiguous() if self.boxpc_one_hot else p_pool
        h = P[4].forward(pf)
        h = P[5].forward(h)
        out9 = P[6].forward(h)
        fit_logits = out9[:, 7:9].contiguous()
        # delta / refine bookkeeping -> F2_* end points (train_semisup_adv.py:368-411)
        E = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        fit_l, fit_p, pred_fit = E(B, 2), E(B), torch.empty((B,), dtype=torch.int32, device=dev)
        d_c, d_s, d_a = E(B, 3), E(B, 3), E(B)
        curr = tuple(t.clone() for t in F_reg)
        tot = (torch.zeros((B, 3), device=dev), torch.zeros((B, 3), device=dev), torch.zeros((B,), device=dev))
        ra = t3d_refine_args(ptr(out9), B, 1 if c.BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF else 0,
                             1 if c.SEMI_WEIGH_BOXPC_DELTA_DURING_TEST else 0, ptr(fit_l), ptr(fit_p), ptr(pred_fit), ptr(d_c),
                             ptr(d_s), ptr(d_a), ptr(curr[0]), ptr(curr[1]), ptr(curr[2]), ptr(tot[0]), ptr(tot[1]), ptr(tot[2]))
        call('t3d_boxpc_refine', ctypes.byref(ra), stream())
        f2c, f2h, f2s = E(B, 3), E(B, NH), E(B, NS, 3)
        call('t3d_f2', ptr(ep['F_center']), ptr(ep['F_heading_residuals']), ptr(ep['F_size_residuals']), ptr(tot[0]), ptr(tot[2]),
             ptr(tot[1]), B, NH, NS, ptr(f2c), ptr(f2h), ptr(f2s), stream())
        ep.update({'boxpc_fit_logits': fit_logits, 'boxpc_fit_prob': fit_p, 'pred_boxpc_fit': pred_fit, 'boxpc_delta_center': d_c,
                   'boxpc_delta_size': d_s, 'boxpc_delta_angle': d_a, 'F2_center': f2c, 'F2_heading_scores': ep['F_heading_scores'],
                   'F2_heading_residuals': f2h, 'F2_size_scores': ep['F_size_scores'], 'F2_size_residuals': f2s})

        # ---- losses + gradients (semisup_v1_sunrgbd.py:323-421), one O(B) kernel + the O(B*N) mask cross-entropy
        res = losses.semi_loss(c, F_output, stage1_center, one_hot, feed, dev, logits=logits, fit_logits=fit_logits, F_reg=F_reg,
                               icv_mask=self.icv_mask, finish=False, mean_size=self.mean_size, orient_anchors=self.orient_anchors)
        dF, ds1, g_reg, dfit, total, per_sample = res['dF'], res['ds1'], res['g_reg'], res['dfit'], res['total'], res['per_sample']

        # ---- backward: BoxPC (input gradient only) -> g_reg
        if float(c.SEMI_WEIGHT_BOXPC_FIT_LOSS) != 0:
            g9 = torch.zeros((B, 9), device=dev)
            g9[:, 7:9] = dfit
            g = P[6].backward(g9)
            g = P[5].backward(g)
            g = P[4].backward(g)
            if self.boxpc_one_hot:
                g = g[:, :512].contiguous()
            # the pooled gradient lives in the <= 512 arg-max rows of each frustum and, through these eval-mode layers (no
            # batch statistics), stays there: the whole chain runs on the compacted rows (B x 512 instead of B x N)
            rows, slot, _, Sr = pool_rows(p_arg, B, N, 512)
            g = scatter_pool_grad(g.contiguous(), slot, B, 512, Sr)
            g = P[3].backward(g, out=gather_rows(P[3].out, rows, B, N, Sr))
            g = P[2].backward(g, out=gather_rows(P[2].out, rows, B, N, Sr))
            g = P[1].backward(g, out=gather_rows(P[1].out, rows, B, N, Sr))
            g6 = P[0].backward(g, k_lo=C, k_hi=C + 6, out=gather_rows(P[0].out, rows, B, N, Sr))     # only the 6 plane-distance channels depend on the box
            pc_rows = gather_rows(pc.reshape(B * N, C), rows, B, N, Sr)
            call('t3d_boxpc_features_bwd', ptr(pc_rows), B, Sr, C, ptr(F_reg[0]), ptr(F_reg[2]), ptr(g6), ptr(g_reg), stream())
            del g, g6
        losses.finish_box_reg(res)

        # ---- backward: refinement head -> feats_lv1
        g = R[2].backward(dF)
        g = dropout(g, m1, keep)
        g = R[1].backward(g)
        g = dropout(g, m0, keep)
        g = R[0].backward(g)
        g_lv1 = g[:, :512].contiguous() if c.use_one_hot else g
        train_box = 'class_agnostic/box_est/conv-reg1/weights' in self.grad
        train_tnet = 'class_agnostic/tnet/fc3-stage1/weights' in self.grad
        if train_box or train_tnet:
            # with the box net frozen (SEMI_TRAIN_BOX_TRAIN_CLASS_AG_BOX = 0) its convolutions still carry the gradient to
            # stage1_center: TrainLayer.backward is dgrad-only for layers outside the gradient arena
            g = Bx[3].backward_pooled(g_lv1.contiguous(), b_arg, B, N, rowmask)
            g = Bx[2].backward(g)
            g = Bx[1].backward(g)
            gx = Bx[0].backward(g, need_dx=train_tnet)
            if train_tnet:
                # the box net sees xyz - stage1_center: d stage1_center -= sum_n dX
                gs = torch.empty((B, 3), device=dev)
                call('t3d_group_sum', ptr(gx), B, N, 3, -1.0, ptr(gs), stream())
                ds1 = ds1 + gs
            del g, gx
        ep['d_stage1_center'], ep['d_feats_lv1'] = ds1, g_lv1
        if train_tnet:
            g = Tn[5].backward(ds1.contiguous().clone())
            g = Tn[4].backward(g)
            g = Tn[3].backward(g)
            g = Tn[2].backward_pooled(g.contiguous(), t_arg, B, N, rowmask)
            g = Tn[1].backward(g)
            Tn[0].backward(g, need_dx=False)
            del g
        ep.update({'semi_loss': total[0:1], 'loss_terms': total, 'per_sample_losses': per_sample, 'mask': mask, 'mask_count': count})
        return ep

    def apply_gradients(self):
        """optimizer.minimize(semi_loss, global_step=batch, var_list=train_vars) (train_semisup_adv.py:420-422)."""
        world = 1
        if self.pg is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            world = torch.distributed.get_world_size(self.pg)
        lr = get_learning_rate(self.global_step, self.B, self.base_lr, self.decay_step, self.decay_rate)
        self.arena.adam_step(lr, self.global_step + 1, world=world, pg=self.pg)
        self.global_step += 1

    def step(self, feed, dropout_masks):
        """sess.run([..., logits, semi_loss, train_semi_op]) of train_one_epoch (train_semisup_adv.py:604-614)."""
        ep = self.forward_backward(feed, dropout_masks)
        self.apply_gradients()
        ep['step'] = self.global_step
        return ep
