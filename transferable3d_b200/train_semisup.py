"""Mirror of the model-A ("backbone") semi-supervised training graph and step of
sunrgbd/sunrgbd_detection/train_semisup.py on the B200:

  get_learning_rate / get_bn_decay (:118-136), train() graph (:199-262): tf_normalize_2D_bboxes -> get_semi_model with
  SEMI_MODEL 'A' (semisup_v1_sunrgbd.get_semi_model_backbone :81-130: inst_seg -> mask / centroid -> T-Net -> box
  estimation net, all in training mode, the one-hot class vector routed into all three) -> get_semi_loss_backbone
  (:256-321: seg cross-entropy + strong box losses on the 3D samples, relaxed reprojection + surface loss on the 2D samples)
  -> optimizer.minimize(semi_loss, global_step=batch) over ALL trainable variables (no var_list), the sess.run of
  train_one_epoch (:330-420).

Unlike the adversarial step (train_semisup_adv.py, BASELINE cfg5) the segmentation network trains here: its backward pass
runs through the folded conv6 (the per-frustum global half receives the sum over the points of the layer's gradient), the
max-pool of conv5 and the two consumers of conv3's point features.  Same kernels as the cfg5 step (lazy batch norm,
max-pool + BN backward without the dense pooled gradient, fused loss kernel) plus t3d_soft_mask / t3d_seg_ce_bwd /
t3d_group_colsum.  Multi-GPU: data parallel, one all-reduce of the flat gradient arena per step.
"""
import numpy as np
import torch

from . import runtime as rt
from . import tf_util, losses, weak_losses
from ._lib import ptr, stream, call
from .constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER, MEAN_DIMS_ARR
from .weights import net_table
from .train_boxpc import get_learning_rate, get_bn_decay          # same schedules (train_semisup.py:118-136)
from .train_layers import ParamArena, TrainLayer, ACT_NONE, ACT_RELU, maxpool, dropout, gemm, dense


class SemiTrainGraph(object):
    def __init__(self, variables, FLAGS, batch_size, num_point, num_channels=6, device='cuda', base_learning_rate=0.001,
                 decay_step=800000, decay_rate=0.5, process_group=None):
        c = FLAGS
        if c.SEMI_MODEL != 'A':
            raise Exception('SemiTrainGraph is the model-A graph of train_semisup.py; SEMI_MODEL %s trains through '
                            'train_semisup_adv.SemiAdvTrainGraph' % c.SEMI_MODEL)
        if c.USE_NORMALIZED_BOX2D_AS_FEATS:
            raise NotImplementedError('USE_NORMALIZED_BOX2D_AS_FEATS in the training graph')
        self.FLAGS, self.B, self.Npt, self.C = FLAGS, batch_size, num_point, num_channels
        self.device = dev = torch.device(device)
        self.base_lr, self.decay_step, self.decay_rate = base_learning_rate, decay_step, decay_rate
        self.pg = process_group
        self.global_step = 0
        self.oh = 10 if c.use_one_hot else 0
        D = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32)).to(dev).contiguous()
        oh = bool(c.use_one_hot)
        nets = (('inst_seg', net_table('inst_seg', num_channels, one_hot=oh)), ('tnet', net_table('tnet', one_hot=oh)),
                ('box_est', net_table('box_est', one_hot=oh)))
        names, self.moving = [], {}
        for scope, table in nets:
            for lname, kind, kw, cin, cout, bn in table:
                layer = '%s/%s' % (scope, lname)
                for suf in ('weights', 'biases') + (('bn/gamma', 'bn/beta') if bn else ()):
                    names.append('%s/%s' % (layer, suf))
                if bn:
                    for suf in ('bn/moving_mean', 'bn/moving_variance'):
                        self.moving['%s/%s' % (layer, suf)] = D(variables['%s/%s' % (layer, suf)])
        self.arena = ParamArena(variables, names, dev)
        self.param, self.grad = self.arena.param, self.arena.grad

        def layers(scope, table):
            return [TrainLayer('%s/%s' % (scope, lname), kw * cin if kind == 'conv' else cin, cout, bn, ACT_RELU if bn else ACT_NONE,
                               self.param, self.moving, self.grad) for lname, kind, kw, cin, cout, bn in table]
        self.seg, self.tnet, self.box = layers(*nets[0]), layers(*nets[1]), layers(*nets[2])
        self.mean_size = D(MEAN_DIMS_ARR)
        self.orient_anchors = D(np.arange(0, 2 * np.pi, 2 * np.pi / NUM_HEADING_BIN))

    def variables(self):
        """Current values keyed by TF variable name (what tf.train.Saver would write)."""
        out = {k: v.detach().cpu().numpy().copy() for k, v in self.param.items()}
        out.update({k: v.detach().cpu().numpy().copy() for k, v in self.moving.items()})
        return out

    def forward_backward(self, feed, dropout_masks):
        """feed: dict keyed like semisup_v1_sunrgbd.placeholder_inputs; dropout_masks: {'inst_seg/dp1': (B,N,128) keep mask}.
        Leaves the gradients in self.grad; returns the loss terms and end points."""
        c, dev = self.FLAGS, self.device
        T = lambda v, dt=torch.float32: (v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v))).to(device=dev, dtype=dt).contiguous()
        B, N, C, OH = self.B, self.Npt, self.C, self.oh
        E = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        pc, one_hot = T(feed['pc']), T(feed['one_hot'])
        assert pc.shape == (B, N, C)
        bn_decay = get_bn_decay(self.global_step, B, self.decay_step)
        cat_oh = lambda t: torch.cat([t, one_hot], dim=1).contiguous() if OH else t
        ep = {}

        # ---- inst_seg in training mode (semisup_models.py:69-139); conv6 = point half (per point) + global half (per frustum)
        S = self.seg
        x = pc.reshape(B * N, C)
        x = S[0].forward(x, bn_decay, lazy=True)
        x = S[1].forward(x, bn_decay, lazy=True)
        pf_lazy = S[2].forward(x, bn_decay, lazy=True)
        x = S[3].forward(pf_lazy, bn_decay, lazy=True)
        x5 = S[4].forward(x, bn_decay, lazy=True)
        gfeat, arg5 = maxpool(x5, B, N, 1024)
        point_feat = dense(pf_lazy)                                       # [B*N, 64]
        G = cat_oh(gfeat)                                                 # [B, 1024 (+10)]: the tiled half of conv6's input
        W6 = S[5].W()                                                     # [64 + 1024 (+10), 512]
        gb = gemm(G, G.shape[1], 1, W6[64:], 512, 1, B, 512, G.shape[1], bias=S[5].p('biases'))
        y6 = E(B * N, 512)
        call('t3d_linear_f32', ptr(point_feat), 64, ptr(W6), 512, None, ptr(gb), N, ptr(y6), 512, B * N, 64, 512, 0, None, None, stream())
        x = S[5].forward(None, bn_decay, y=y6, lazy=True)
        x = S[6].forward(x, bn_decay, lazy=True)
        x = S[7].forward(x, bn_decay, lazy=True)
        x = S[8].forward(x, bn_decay)
        keep = T(dropout_masks['inst_seg/dp1']).reshape(B * N, 128)
        x = dropout(x, keep, 0.5)
        logits = S[9].forward(x, bn_decay).reshape(B, N, 2)
        ep['logits'] = logits
        soft_mask = E(B, N)
        call('t3d_soft_mask', ptr(logits), B, N, ptr(soft_mask), stream())
        ep['soft_mask'] = soft_mask

        # ---- mask, centroid (semisup_models.py:145-162): a compare, no gradient
        mask, count, mean, xyz1, _ = rt.mask_centroid(logits, pc, want_mask=True, want_xyz_stage1=True, want_idx=False)
        rowmask = mask.reshape(B * N)

        # ---- T-Net (semisup_models.py:164-202)
        Tn = self.tnet
        x = xyz1.reshape(B * N, 3)
        for l in Tn[:3]:
            x = l.forward(x, bn_decay, lazy=True)
        t_pool, t_arg = maxpool(x, B, N, 256, rowmask)
        h = Tn[3].forward(cat_oh(t_pool), bn_decay)
        h = Tn[4].forward(h, bn_decay)
        t_out = Tn[5].forward(h, bn_decay)
        stage1_center = (t_out + mean).contiguous()
        ep['stage1_center'] = stage1_center

        # ---- box estimation net (semisup_models.py:204-291), the prediction head of model A
        Bx = self.box
        xin = E(B, N, 3)
        call('t3d_prepare_xyz', ptr(pc), B, N, C, ptr(stage1_center), ptr(xin), stream())
        x = xin.reshape(B * N, 3)
        for l in Bx[:4]:
            x = l.forward(x, bn_decay, lazy=True)
        feats_lv1, b_arg = maxpool(x, B, N, 512, rowmask)
        h = Bx[4].forward(cat_oh(feats_lv1), bn_decay)
        h = Bx[5].forward(h, bn_decay)
        box_params = Bx[6].forward(h, bn_decay)
        ep['feats_lv1'], ep['box_params'] = feats_lv1, box_params
        bp = tf_util.parse_box_output(box_params, stage1_center, self.mean_size, self.orient_anchors, want_reg=True)
        for k in ('center', 'heading_scores', 'heading_residuals_normalized', 'heading_residuals', 'size_scores',
                  'size_residuals_normalized', 'size_residuals'):
            ep[k] = bp[k]
        S_reg = bp['reg']
        ep['S_pred_box_reg'] = S_reg

        # ---- get_semi_loss_backbone (semisup_v1_sunrgbd.py:256-321): fused loss kernel (strong + reprojection) + surface kernel
        res = losses.semi_loss(c, box_params, stage1_center, one_hot, feed, dev, logits=logits, F_reg=S_reg, finish=False,
                               mean_size=self.mean_size, orient_anchors=self.orient_anchors, model_a=True)
        total = res['total']
        is2d = T(feed['is_data_2D'])
        gmask = None
        if float(c.WEAK_WEIGHT_SURFACE) != 0.0:
            up = (is2d * (float(c.SEMI_MULTIPLIER_FOR_WEAK_LOSS) * float(c.WEAK_WEIGHT_SURFACE) / B)).contiguous()
            sl = weak_losses.get_surface_loss(S_reg, pc, soft_mask, margin=c.WEAK_SURFACE_MARGIN,
                                              scale_dims_factor=c.WEAK_SURFACE_LOSS_SCALE_DIMS,
                                              weight_for_points_within=c.WEAK_SURFACE_LOSS_WT_FOR_INNER_PTS,
                                              train_seg=c.WEAK_TRAIN_SEG_W_SURFACE, train_box=c.WEAK_TRAIN_BOX_W_SURFACE,
                                              reduce_loss=False, end_points=ep, upstream=up)
            res['g_reg'].add_(ep['surface_grad_box_reg'])
            gmask = ep['surface_grad_soft_mask']
            total = total.clone()
            total[0] += (sl * up).sum()
        losses.finish_box_reg(res)
        dF, ds1 = res['dF'], res['ds1']
        wb = ((1.0 - is2d) * (float(c.STRONG_WEIGHT_CROSS_ENTROPY) / B)).contiguous()
        dlogits = E(B, N, 2)
        call('t3d_seg_ce_bwd', ptr(logits), ptr(T(feed['labels'], torch.int32)), ptr(wb), ptr(gmask), B, N, ptr(dlogits), stream())

        # ---- backward: box head -> box convs -> stage1_center -> T-Net
        g = Bx[6].backward(dF)
        g = Bx[5].backward(g)
        g = Bx[4].backward(g)
        g_lv1 = g[:, :512].contiguous() if OH else g
        g = Bx[3].backward_pooled(g_lv1, b_arg, B, N, rowmask)
        g = Bx[2].backward(g)
        g = Bx[1].backward(g)
        gx = Bx[0].backward(g)
        gs = E(B, 3)
        call('t3d_group_sum', ptr(gx), B, N, 3, -1.0, ptr(gs), stream())        # the box net sees xyz - stage1_center
        ds1 = (ds1 + gs).contiguous()
        g = Tn[5].backward(ds1.clone())
        g = Tn[4].backward(g)
        g = Tn[3].backward(g)
        g_t = g[:, :256].contiguous() if OH else g
        g = Tn[2].backward_pooled(g_t, t_arg, B, N, rowmask)
        g = Tn[1].backward(g)
        Tn[0].backward(g, need_dx=False)

        # ---- backward: segmentation net
        g = S[9].backward(dlogits.reshape(B * N, 2))
        g = dropout(g, keep, 0.5)
        g = S[8].backward(g)
        g = S[7].backward(g)
        g = S[6].backward(g)
        dY6 = S[5].backward_bn_only(g)                                    # [B*N, 512]
        dW6 = self.grad['inst_seg/conv6/weights'].view(-1, 512)
        self.grad['inst_seg/conv6/biases'].zero_()                        # a bias in front of a batch norm: zero gradient
        from .train_layers import splitk_for
        call('t3d_gemm_f32', ptr(point_feat), 1, 64, ptr(dY6), 512, 1, ptr(dW6[:64]), 512, 64, 512, B * N, splitk_for(64, 512, B * N),
             None, stream())
        Ssum = E(B, 512)
        call('t3d_group_colsum', ptr(dY6), B, N, 512, ptr(Ssum), stream())
        KG = G.shape[1]
        call('t3d_gemm_f32', ptr(G), 1, KG, ptr(Ssum), 512, 1, ptr(dW6[64:]), 512, KG, 512, B, 1, None, stream())
        dG = gemm(Ssum, 512, 1, W6[64:], 1, 512, B, KG, 512)             # [B, 1024 (+10)]
        dgfeat = dG[:, :1024].contiguous()
        dpf = gemm(dY6, 512, 1, W6[:64], 1, 512, B * N, 64, 512)         # [B*N, 64]
        g = S[4].backward_pooled(dgfeat, arg5, B, N)
        g = S[3].backward(g)
        dpf.add_(g)                                                       # conv3's point features feed conv4 and conv6
        g = S[2].backward(dpf)
        g = S[1].backward(g)
        S[0].backward(g, need_dx=False)
        ep.update({'semi_loss': total[0:1], 'loss_terms': total, 'per_sample_losses': res['per_sample'], 'mask': mask, 'mask_count': count})
        return ep

    def apply_gradients(self):
        """optimizer.minimize(semi_loss, global_step=batch) (train_semisup.py:243-244)."""
        world = 1
        if self.pg is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            world = torch.distributed.get_world_size(self.pg)
        lr = get_learning_rate(self.global_step, self.B, self.base_lr, self.decay_step, self.decay_rate)
        self.arena.adam_step(lr, self.global_step + 1, world=world, pg=self.pg)
        self.global_step += 1

    def step(self, feed, dropout_masks):
        ep = self.forward_backward(feed, dropout_masks)
        self.apply_gradients()
        ep['step'] = self.global_step
        return ep
