"""Mirror of sunrgbd/sunrgbd_detection/semisup_models.py (same function names, argument order and
end_points keys), evaluated on the B200.

bf16 / f16x2 modes: every per-point conv stack + max-pool is ONE fused tcgen05 kernel (csrc/chain_max.cuh,
csrc/seg_stage2_pair.cuh; split-precision twins csrc/chain_x2.cuh, csrc/seg_stage2_x2.cuh); the masked stacks (tnet,
box_est) run on the compacted masked-in points, which equals the reference's max(net*mask) because post-ReLU
activations are >= 0.
fp32 mode: layer-by-layer fp32 GEMMs (t3d_linear_f32) with the literal mask multiply.
Variables are resolved by TF name through runtime.variable_scope (checkpoint contract).
"""
import numpy as np
import torch

from . import runtime as rt
from . import tf_util
from ._lib import ptr, stream, call
from .constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER, MEAN_DIMS_ARR, ORIENT_ANCHORS


class PointView(object):
    """Lazy (B,N,3) tensor  pc[:,:,0:3] - center[:,None,:]  (the reference materialises these in
    subtract_points_mean / subtract_1st_stage_center; the fused kernels subtract on load)."""

    def __init__(self, pc, center=None):
        self.pc = pc
        self.center = None if center is None else rt.f32(center.reshape(-1, 3))

    @property
    def shape(self):
        return (self.pc.shape[0], self.pc.shape[1], 3)

    def tensor(self):
        B, N, C = self.pc.shape
        out = torch.empty((B, N, 3), dtype=torch.float32, device=self.pc.device)
        call('t3d_prepare_xyz', ptr(self.pc), B, N, C, ptr(self.center), ptr(out), stream())
        return out


def _as_view(x):
    if isinstance(x, PointView):
        return x
    return PointView(rt.f32(x), None)


def _compaction_of(mask, pc):
    """(idx, count) of a (B,N,1)/(B,N) 0/1 mask; reuses what subtract_points_mean computed."""
    if hasattr(mask, '_t3d_idx'):
        return mask._t3d_idx, mask._t3d_count
    m = rt.f32(mask.reshape(mask.shape[0], mask.shape[1]))
    fake_logits = torch.stack([torch.full_like(m, 0.5), m], dim=2).contiguous()
    _, count, _, _, idx = rt.mask_centroid(fake_logits, pc, want_mask=False)
    return idx, count


def _cat(parts):
    parts = [rt.f32(p) for p in parts if p is not None]
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1).contiguous()


def mlps(input_feat, layers, is_training, bn=True, bn_decay=None, c=None, scope=None, reuse=None):
    """semisup_models.py:30-42: fully connected stack; the last layer has neither batch norm nor activation."""
    rt.require_eval(is_training)
    with rt.variable_scope(scope):
        net = input_feat
        for i, layer_size in enumerate(layers):
            last = (len(layers) - 1 == i)
            net = tf_util.fully_connected(net, layer_size, scope='fc%d' % i, bn=(False if last else bn),
                                          activation_fn=(None if last else 'relu'), is_training=is_training, bn_decay=bn_decay)
    return net


def _normalize_xyz(pc, normalize_method):
    """semisup_models.py:335-343 / :413-421: xyz normalised per cloud, the remaining channels untouched (t3d_normalize_pc)."""
    if normalize_method == 'SD':
        return tf_util.tf_normalize_point_clouds_to_mean_zero_and_unit_var(pc)
    if normalize_method == 'Spread':
        return tf_util.tf_normalize_point_clouds_to_01(pc)
    raise Exception('Invalid normalization method')


def _conv_stack_max(x3d, names, full, mask=None):
    """conv2d + BN + ReLU stack over (B,N,D) points followed by (* mask and) the max over points, layer by layer through
    t3d_linear_f32 (BN folded): the path of the inputs the fused chains do not take (normalised clouds, an extra mask
    channel) and of the fp32 mode."""
    st = rt.store()
    B, N, D = x3d.shape
    x = rt.f32(x3d).reshape(B * N, D)
    for name in names[:-1]:
        w, b = st.folded(full + '/' + name)
        x, _ = rt.linear(x, w, b, 'relu')
    w, b = st.folded(full + '/' + names[-1])
    rowmask = None if mask is None else rt.f32(mask).reshape(B * N).contiguous()
    _, net = rt.linear(x, w, b, 'relu', rows_per_group=N, gmax_groups=B, want_y=False, rowmask=rowmask)
    return net


def mlps_with_dropout(input_feat, layers, activation_fns, keep_probs, is_training, bn=True, bn_decay=None,
                      c=None, scope=None, reuse=None):
    """semisup_models.py:44-63 (activation_fns are 'relu' / 'leaky_relu' / 'tanh' / None)."""
    assert len(layers) == len(activation_fns) == len(keep_probs)
    rt.require_eval(is_training)
    with rt.variable_scope(scope):
        net = input_feat
        for i, layer_size in enumerate(layers):
            last = (len(layers) - 1 == i)
            net = tf_util.fully_connected(net, layer_size, scope='fc%d' % i, activation_fn=activation_fns[i],
                                          is_training=is_training, bn=(False if last else bn), bn_decay=bn_decay)
            if not last:
                net = tf_util.dropout(net, keep_prob=keep_probs[i], is_training=is_training, scope='dp%d' % i)
    return net


def v1_inst_seg(point_cloud, img_feats, one_hot_vec, end_points, is_training, bn_decay=None, scope=None):
    """semisup_models.py:69-139: (B,N,D) -> mask logits (B,N,2)."""
    rt.require_eval(is_training)
    st = rt.store()
    pc = rt.f32(point_cloud)
    B, N, D = pc.shape
    with rt.variable_scope(scope):
        full = st.scope_name()
        w6, b6 = st.folded(full + '/conv6')
        if rt.get_precision() in rt.FUSED:
            if D != 6:
                raise ValueError('the tcgen05 inst_seg kernel is built for 6-channel frustums (got %d)' % D)
            x2 = rt.is_x2()
            arena1 = st.chain_arena(full, rt.CHAIN_SEG1, ['conv1', 'conv2', 'conv3', 'conv4', 'conv5'], x2=x2)
            arena2 = st.seg2_arena(full, x2=x2)
            # stage 1 emits conv3's output as the swizzled operand images of stage 2 (bf16: one [256 x 64] image per
            # 256-point tile; f16x2: a hi and a lo fp16 [128 x 64] image per 128-point tile)
            if x2:
                point_feat = torch.empty((B * ((N + 127) // 128), 32768), dtype=torch.uint8, device=pc.device)
                w6g, b6g = st.cached(full + '/conv6/x2_global', lambda: ((w6[64:] * rt.X2_ACT_SCALE).contiguous(),
                                                                         (b6 * rt.X2_ACT_SCALE).contiguous()))
            else:
                point_feat = torch.empty((B * ((N + 255) // 256) * 256, 64), dtype=torch.bfloat16, device=pc.device)
                w6g, b6g = st.cached(full + '/conv6/global', lambda: (w6[64:].contiguous(), b6))
            gfeat = rt.chain_max(rt.CHAIN_SEG1, pc, arena1, emit=point_feat, x2=x2)          # (B,1024)
            g = _cat([gfeat, one_hot_vec])
            gbias, _ = rt.linear(g, w6g, b6g)                                               # conv6 global half
            logits = rt.seg_stage2(point_feat, gbias, arena2, B, N, x2=x2)
        else:
            if isinstance(pc, rt.WirePoints):
                pc = pc.dense()
            x = pc.reshape(B * N, D)
            for name in ('conv1', 'conv2', 'conv3'):
                w, b = st.folded(full + '/' + name)
                x, _ = rt.linear(x, w, b, 'relu')
            point_feat = x
            w, b = st.folded(full + '/conv4')
            x, _ = rt.linear(point_feat, w, b, 'relu')
            w, b = st.folded(full + '/conv5')
            _, gfeat = rt.linear(x, w, b, 'relu', rows_per_group=N, gmax_groups=B, want_y=False)
            g = _cat([gfeat, one_hot_vec])
            gbias, _ = rt.linear(g, w6[64:].contiguous(), b6)
            x, _ = rt.linear(point_feat, w6[:64].contiguous(), None, 'relu', gbias=gbias, rows_per_group=N)
            for name in ('conv7', 'conv8', 'conv9'):
                w, b = st.folded(full + '/' + name)
                x, _ = rt.linear(x, w, b, 'relu')
            w, b = st.folded(full + '/conv10')
            x, _ = rt.linear(x, w, b, None)
            logits = x.reshape(B, N, 2)
    return logits


def subtract_points_mean(point_cloud, logits, scope=None):
    """semisup_models.py:145-162 -> (mask (B,N,1), mask_xyz_mean (B,1,3), xyz, xyz_stage1).
    xyz / xyz_stage1 are PointViews (call .tensor() for the dense (B,N,3) tensor)."""
    pc = rt.f32(point_cloud)
    mask, count, mean, _, idx = rt.mask_centroid(logits, pc)
    mask3 = mask.unsqueeze(2)
    mask3._t3d_idx, mask3._t3d_count = idx, count
    return mask3, mean.unsqueeze(1), PointView(pc, None), PointView(pc, mean)


def _masked_chain(kind, view, mask, layer_names, scope_full):
    """conv stack on a PointView + mask + max over points -> (B, Cout)."""
    st = rt.store()
    pc = view.pc
    B, N, _ = pc.shape
    if rt.get_precision() in rt.FUSED:
        x2 = rt.is_x2()
        arena = st.chain_arena(scope_full, kind, layer_names, x2=x2)
        if mask is None:
            return rt.chain_max(kind, pc, arena, center=view.center, x2=x2)
        idx, count = _compaction_of(mask, pc)
        return rt.chain_max(kind, pc, arena, center=view.center, idx=idx, count=count, x2=x2)
    x = view.tensor().reshape(B * N, 3)
    rowmask = None if mask is None else rt.f32(mask.reshape(B * N))
    for name in layer_names[:-1]:
        w, b = st.folded(scope_full + '/' + name)
        x, _ = rt.linear(x, w, b, 'relu')
    w, b = st.folded(scope_full + '/' + layer_names[-1])
    _, feat = rt.linear(x, w, b, 'relu', rows_per_group=N, rowmask=rowmask, gmax_groups=B, want_y=False)
    return feat


def v1_tnet(point_cloud_xyz_stage1, mask, mask_xyz_mean, one_hot_vec, end_points, is_training, norm_box2D=None,
            bn_decay=None, scope=None):
    """semisup_models.py:164-202 -> stage1_center (B,3)."""
    rt.require_eval(is_training)
    view = _as_view(point_cloud_xyz_stage1)
    with rt.variable_scope(scope):
        full = rt.store().scope_name()
        net = _masked_chain(rt.CHAIN_TNET, view, mask, ['conv-reg1-stage1', 'conv-reg2-stage1', 'conv-reg3-stage1'], full)
        net = _cat([net, one_hot_vec, norm_box2D])
        net = tf_util.fully_connected(net, 256, scope='fc1-stage1', bn=True, is_training=is_training, bn_decay=bn_decay)
        net = tf_util.fully_connected(net, 128, scope='fc2-stage1', bn=True, is_training=is_training, bn_decay=bn_decay)
        stage1_center = tf_util.fully_connected(net, 3, activation_fn=None, scope='fc3-stage1')
        stage1_center = stage1_center + rt.f32(mask_xyz_mean).reshape(-1, 3)
        end_points['stage1_center'] = stage1_center
        return stage1_center


def subtract_1st_stage_center(point_cloud_xyz, stage1_center, scope=None):
    """semisup_models.py:204-209: original xyz - stage1_center (lazy)."""
    v = _as_view(point_cloud_xyz)
    if v.center is not None:
        return PointView(v.pc, v.center + rt.f32(stage1_center))
    return PointView(v.pc, stage1_center)


def parse_into_end_points(output, stage1_center, end_points, prefix):
    """semisup_models.py:265-290 (and the copy semisup_v1_sunrgbd.py:203-222)."""
    st = rt.store()
    p = tf_util.parse_box_output(output, rt.f32(stage1_center), st.const('MEAN_DIMS_ARR', MEAN_DIMS_ARR),
                                 st.const('ORIENT_ANCHORS', ORIENT_ANCHORS))
    for k in ('center', 'heading_scores', 'heading_residuals_normalized', 'heading_residuals', 'size_scores',
              'size_residuals_normalized', 'size_residuals'):
        end_points[prefix + k] = p[k]
    pred_box = (p['center'], p['size_scores'], p['size_residuals'], p['heading_scores'], p['heading_residuals'])
    return pred_box, p['reg']


def v1_box_est(point_cloud_xyz_submean, stage1_center, mask, one_hot_vec, end_points, is_training, norm_box2D=None,
               bn_decay=None, prefix='', c=None, scope=None):
    """semisup_models.py:215-291 -> pred_box = (center, size_scores, size_residuals, heading_scores, heading_residuals)."""
    rt.require_eval(is_training)
    view = _as_view(point_cloud_xyz_submean)
    with rt.variable_scope(scope):
        full = rt.store().scope_name()
        net = _masked_chain(rt.CHAIN_BOX, view, mask, ['conv-reg1', 'conv-reg2', 'conv-reg3', 'conv-reg4'], full)
        end_points[prefix + 'feats_lv1'] = net
        net = _cat([net, one_hot_vec, norm_box2D])
        net = tf_util.fully_connected(net, 512, scope='fc1', bn=True, is_training=is_training, bn_decay=bn_decay)
        end_points[prefix + 'feats_lv2'] = net
        net = tf_util.fully_connected(net, 256, scope='fc2', bn=True, is_training=is_training, bn_decay=bn_decay)
        end_points[prefix + 'feats_lv3'] = net
        output = tf_util.fully_connected(net, 3 + NUM_HEADING_BIN * 2 + NUM_SIZE_CLUSTER * 4, activation_fn=None, scope='fc3')
        end_points[prefix + 'box_params'] = output
        pred_box, reg = parse_into_end_points(output, stage1_center, end_points, prefix)
        end_points[prefix + '_box_reg_fused'] = reg
    return pred_box


def box_pc_mask_features_model(box, pc, logits, num_outputs, is_training, end_points, reuse, bn_for_output,
                               normalize_pc=False, normalize_method='SD', one_hot_vec=None, norm_box2D=None,
                               bn_decay=None, c=None, scope=None):
    """semisup_models.py:297-324."""
    if c.BOX_PC_MASK_REPRESENTATION == 'A':
        fn = combined_box_pc_mask_features_model
    elif c.BOX_PC_MASK_REPRESENTATION == 'B':
        fn = independent_box_pc_mask_features_model
    else:
        raise Exception('Box pc mask representation not implemented: %s' % c.BOX_PC_MASK_REPRESENTATION)
    return fn(box, pc, logits, num_outputs, is_training, end_points=end_points, reuse=reuse, normalize_pc=normalize_pc,
              normalize_method=normalize_method, bn_for_output=False, one_hot_vec=one_hot_vec, norm_box2D=None,
              bn_decay=bn_decay, c=c, scope='box_pc_mask_model')


def combined_box_pc_mask_features_model(box_reg, pc, mask, num_outputs, is_training, end_points, reuse, bn_for_output,
                                        normalize_pc=False, normalize_method='SD', one_hot_vec=None, norm_box2D=None,
                                        bn_decay=None, c=None, scope=None):
    """semisup_models.py:326-398.  mask=None, normalize_pc=False (what every caller in the reference passes) runs the fused
    chain with the plane distances in its prologue; a normalised representation or a mask channel takes the layer-wise
    path on the materialised (B,N,C+6[+1]) representation."""
    if normalize_pc and normalize_method not in ('SD', 'Spread'):
        raise Exception('Invalid normalization method')        # semisup_models.py:343, before anything is built
    rt.require_eval(is_training)
    st = rt.store()
    pc = rt.f32(pc)
    B, N, C = pc.shape
    names = ['conv-reg1', 'conv-reg2', 'conv-reg3', 'conv-reg4']
    with rt.variable_scope(scope):
        full = st.scope_name()
        if not normalize_pc and mask is None and rt.get_precision() in rt.FUSED:
            arena = st.chain_arena(full, rt.CHAIN_BOXPC, names, x2=rt.is_x2())
            net = rt.chain_max(rt.CHAIN_BOXPC, pc, arena, box=box_reg, x2=rt.is_x2())
        else:
            rep = tf_util.tf_get_box_pc_representation(box_reg, pc)
            if normalize_pc:
                rep = _normalize_xyz(rep, normalize_method)
            if mask is not None:
                rep = torch.cat([rep, rt.f32(mask).reshape(B, N, 1)], dim=2).contiguous()
            net = _conv_stack_max(rep, names, full, mask)
        net = _cat([net, one_hot_vec, norm_box2D])
        features_lv1 = net
        net = tf_util.fully_connected(net, 512, bn=True, is_training=is_training, scope='fc1', bn_decay=bn_decay)
        features_lv2 = net
        net = tf_util.dropout(net, keep_prob=0.7, is_training=is_training, scope='dp1')
        net = tf_util.fully_connected(net, 256, bn=True, is_training=is_training, scope='fc2', bn_decay=bn_decay)
        features_lv3 = net
        net = tf_util.dropout(net, keep_prob=0.7, is_training=is_training, scope='dp2')
        net = tf_util.fully_connected(net, num_outputs, bn=bn_for_output, is_training=is_training, activation_fn=None,
                                      scope='fc3', bn_decay=bn_decay)
        features = {'%s_feats_lv1' % scope: features_lv1, '%s_feats_lv2' % scope: features_lv2,
                    '%s_feats_lv3' % scope: features_lv3}
    return net, features


def independent_box_pc_mask_features_model(box_reg, pc, mask, num_outputs, is_training, end_points, reuse, bn_for_output,
                                           normalize_pc=False, normalize_method='SD', one_hot_vec=None, norm_box2D=None,
                                           bn_decay=None, c=None, scope=None):
    """semisup_models.py:400-471 (BoxPC representation B): the box (B,7) goes through the `extract_box_feats` FC stack, the
    raw points through conv 128-128-256-512 + max (one fused chain, CHAIN_BOXPCB, in the bf16 / f16x2 modes), and the
    concatenation [box_feat, point_feat (, norm_box2D) (, one_hot)] through the 4-layer FC head."""
    if normalize_pc and normalize_method not in ('SD', 'Spread'):
        raise Exception('Invalid normalization method')        # semisup_models.py:421
    rt.require_eval(is_training)
    st = rt.store()
    pc = rt.f32(pc)
    B, N, D = pc.shape
    names = ['conv-reg1', 'conv-reg2', 'conv-reg3', 'conv-reg4']
    with rt.variable_scope(scope):
        full = st.scope_name()
        box7 = torch.cat([rt.f32(box_reg[0]).reshape(B, 3), rt.f32(box_reg[1]).reshape(B, 3), rt.f32(box_reg[2]).reshape(B, 1)],
                         dim=1).contiguous()
        box_feat = mlps(box7, [128, 128, 256, 512], is_training, bn=True, bn_decay=bn_decay, c=c, scope='extract_box_feats',
                        reuse=reuse)
        if normalize_pc:
            pc = _normalize_xyz(pc, normalize_method)
        if mask is None and D == 6 and rt.get_precision() in rt.FUSED:
            arena = st.chain_arena(full, rt.CHAIN_BOXPCB, names, x2=rt.is_x2())
            net = rt.chain_max(rt.CHAIN_BOXPCB, pc, arena, x2=rt.is_x2())
        else:
            net = _conv_stack_max(pc, names, full, mask)
        net = _cat([box_feat, net])
        features_lv1 = net
        net = _cat([net, norm_box2D, one_hot_vec])
        net = tf_util.fully_connected(net, 512, bn=True, is_training=is_training, scope='fc1', bn_decay=bn_decay)
        net = tf_util.fully_connected(net, 512, bn=True, is_training=is_training, scope='fc2', bn_decay=bn_decay)
        features_lv2 = net
        net = tf_util.dropout(net, keep_prob=0.7, is_training=is_training, scope='dp2')
        net = tf_util.fully_connected(net, 256, bn=True, is_training=is_training, scope='fc3', bn_decay=bn_decay)
        features_lv3 = net
        net = tf_util.dropout(net, keep_prob=0.7, is_training=is_training, scope='dp3')
        net = tf_util.fully_connected(net, num_outputs, bn=bn_for_output, is_training=is_training, activation_fn=None,
                                      scope='fc4', bn_decay=bn_decay)
        features = {'%s_feats_lv1' % scope: features_lv1, '%s_feats_lv2' % scope: features_lv2,
                    '%s_feats_lv3' % scope: features_lv3}
    return net, features
