"""Sub-networks: restatement of sunrgbd/sunrgbd_detection/semisup_models.py:30-398.

`vs` (tf_layers.VarStore) stands in for the TF variable collection; `scope` nests exactly like
tf.variable_scope so variables resolve to the reference checkpoint names.
"""
import numpy as np
import torch

from . import tf_util
from .tf_layers import conv2d, fully_connected, max_pool_points, dropout
from transferable3d_b200.constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER, MEAN_DIMS_ARR


def mlps(input_feat, layers, is_training, vs, bn=True, bn_decay=None, c=None, scope=None):
    """semisup_models.py:30-42: fully connected stack, the last layer without batch norm and without activation."""
    with vs.variable_scope(scope):
        net = input_feat
        for i, layer_size in enumerate(layers):
            if len(layers) - 1 == i:
                net = fully_connected(net, layer_size, vs, 'fc%d' % i, bn=False, is_training=is_training, activation_fn=None,
                                      bn_decay=bn_decay)
            else:
                net = fully_connected(net, layer_size, vs, 'fc%d' % i, bn=bn, is_training=is_training, bn_decay=bn_decay)
    return net


def _normalize_xyz(pc, normalize_method):
    """semisup_models.py:335-343 / :413-421: xyz normalised, the remaining channels untouched."""
    xyz, rest = pc[:, :, :3], pc[:, :, 3:]
    if normalize_method == 'SD':
        xyz = tf_util.tf_normalize_point_clouds_to_mean_zero_and_unit_var(xyz)
    elif normalize_method == 'Spread':
        xyz = tf_util.tf_normalize_point_clouds_to_01(xyz)
    else:
        raise Exception('Invalid normalization method')
    return torch.cat([xyz, rest], dim=2)


def mlps_with_dropout(input_feat, layers, activation_fns, keep_probs, is_training, vs, bn=True,
                      bn_decay=None, c=None, scope=None):
    """semisup_models.py:44-63."""
    assert len(layers) == len(activation_fns) == len(keep_probs)
    with vs.variable_scope(scope):
        net = input_feat
        for i, layer_size in enumerate(layers):
            if len(layers) - 1 == i:
                net = fully_connected(net, layer_size, vs, 'fc%d' % i, bn=False, is_training=is_training,
                                      activation_fn=activation_fns[i], bn_decay=bn_decay)
            else:
                net = fully_connected(net, layer_size, vs, 'fc%d' % i, bn=bn, is_training=is_training,
                                      activation_fn=activation_fns[i], bn_decay=bn_decay)
                net = dropout(net, vs, is_training, 'dp%d' % i, keep_prob=keep_probs[i])
    return net


def v1_inst_seg(point_cloud, img_feats, one_hot_vec, end_points, is_training, vs, bn_decay=None, scope=None):
    """semisup_models.py:69-139. Literal graph: the global feature is tiled to every point and
    conv6 runs on the 1088(+10)-wide concat (vs.literal=True); vs.literal=False evaluates the
    algebraically identical folded form (global half of conv6 applied once per frustum)."""
    with vs.variable_scope(scope):
        B, N, D = point_cloud.shape
        net = conv2d(point_cloud, 64, [1, D], vs, 'conv1', True, is_training, bn_decay=bn_decay)
        net = conv2d(net, 64, [1, 1], vs, 'conv2', True, is_training, bn_decay=bn_decay)
        point_feat = conv2d(net, 64, [1, 1], vs, 'conv3', True, is_training, bn_decay=bn_decay)
        net = conv2d(point_feat, 128, [1, 1], vs, 'conv4', True, is_training, bn_decay=bn_decay)
        net = conv2d(net, 1024, [1, 1], vs, 'conv5', True, is_training, bn_decay=bn_decay)
        global_feat = max_pool_points(net)                                   # (B,1024)
        if one_hot_vec is not None:
            global_feat = torch.cat([global_feat, one_hot_vec], dim=1)
        end_points['_oracle_global_feat'] = global_feat
        if vs.literal or is_training:
            global_feat_expand = global_feat.unsqueeze(1).repeat(1, N, 1)
            concat_feat = torch.cat([point_feat, global_feat_expand], dim=2)
            net = conv2d(concat_feat, 512, [1, 1], vs, 'conv6', True, is_training, bn_decay=bn_decay)
        else:
            with vs.variable_scope('conv6'):
                w = vs.get('weights').reshape(-1, 512)
                y = point_feat @ w[:64] + (global_feat @ w[64:]).unsqueeze(1) + vs.get('biases')
                from .tf_layers import batch_norm
                net = torch.relu(batch_norm(y, vs, 'bn', False, bn_decay))
        net = conv2d(net, 256, [1, 1], vs, 'conv7', True, is_training, bn_decay=bn_decay)
        net = conv2d(net, 128, [1, 1], vs, 'conv8', True, is_training, bn_decay=bn_decay)
        net = conv2d(net, 128, [1, 1], vs, 'conv9', True, is_training, bn_decay=bn_decay)
        net = dropout(net, vs, is_training, 'dp1', keep_prob=0.5)
        logits = conv2d(net, 2, [1, 1], vs, 'conv10', activation_fn=None)
    return logits


def subtract_points_mean(point_cloud, logits, scope=None):
    """semisup_models.py:145-162."""
    N = point_cloud.shape[1]
    mask = (logits[:, :, 0:1] < logits[:, :, 1:2]).to(point_cloud.dtype)     # (B,N,1) strict <
    mask_count = mask.sum(dim=1, keepdim=True).repeat(1, 1, 3)
    xyz = point_cloud[:, :, 0:3]
    mean = (mask.repeat(1, 1, 3) * xyz).sum(dim=1, keepdim=True)
    mean = mean / torch.clamp(mask_count, min=1)
    xyz_stage1 = xyz - mean.repeat(1, N, 1)
    return mask, mean, xyz, xyz_stage1


def v1_tnet(point_cloud_xyz_stage1, mask, mask_xyz_mean, one_hot_vec, end_points, is_training, vs,
            norm_box2D=None, bn_decay=None, scope=None):
    """semisup_models.py:164-202."""
    with vs.variable_scope(scope):
        net = conv2d(point_cloud_xyz_stage1, 128, [1, 1], vs, 'conv-reg1-stage1', True, is_training, bn_decay=bn_decay)
        net = conv2d(net, 128, [1, 1], vs, 'conv-reg2-stage1', True, is_training, bn_decay=bn_decay)
        net = conv2d(net, 256, [1, 1], vs, 'conv-reg3-stage1', True, is_training, bn_decay=bn_decay)
        masked_net = net * mask                                              # (B,N,256)*(B,N,1)
        net = max_pool_points(masked_net)
        if one_hot_vec is not None:
            net = torch.cat([net, one_hot_vec], dim=1)
        if norm_box2D is not None:
            net = torch.cat([net, norm_box2D], dim=1)
        net = fully_connected(net, 256, vs, 'fc1-stage1', True, is_training, bn_decay=bn_decay)
        net = fully_connected(net, 128, vs, 'fc2-stage1', True, is_training, bn_decay=bn_decay)
        stage1_center = fully_connected(net, 3, vs, 'fc3-stage1', activation_fn=None)
        stage1_center = stage1_center + mask_xyz_mean.squeeze(1)
        end_points['stage1_center'] = stage1_center
        return stage1_center


def subtract_1st_stage_center(point_cloud_xyz, stage1_center, scope=None):
    """semisup_models.py:204-209."""
    return point_cloud_xyz - stage1_center.unsqueeze(1)


def parse_box_output(output, stage1_center, end_points, prefix):
    """The slicing of semisup_models.py:265-290 (and its copy semisup_v1_sunrgbd.py:203-222)."""
    NH, NS = NUM_HEADING_BIN, NUM_SIZE_CLUSTER
    B = output.shape[0]
    center = output[:, 0:3] + stage1_center
    end_points[prefix + 'center'] = center
    end_points[prefix + 'heading_scores'] = output[:, 3:3 + NH]
    hrn = output[:, 3 + NH:3 + 2 * NH]
    end_points[prefix + 'heading_residuals_normalized'] = hrn
    end_points[prefix + 'heading_residuals'] = hrn * (np.pi / NH)
    end_points[prefix + 'size_scores'] = output[:, 3 + 2 * NH:3 + 2 * NH + NS]
    srn = output[:, 3 + 2 * NH + NS:3 + 2 * NH + 4 * NS].reshape(B, NS, 3)
    end_points[prefix + 'size_residuals_normalized'] = srn
    end_points[prefix + 'size_residuals'] = srn * torch.as_tensor(MEAN_DIMS_ARR, dtype=torch.float32).to(output.dtype).unsqueeze(0)
    return (center, end_points[prefix + 'size_scores'], end_points[prefix + 'size_residuals'],
            end_points[prefix + 'heading_scores'], end_points[prefix + 'heading_residuals'])


def v1_box_est(point_cloud_xyz_submean, stage1_center, mask, one_hot_vec, end_points, is_training, vs,
               norm_box2D=None, bn_decay=None, prefix='', c=None, scope=None):
    """semisup_models.py:215-291."""
    with vs.variable_scope(scope):
        net = conv2d(point_cloud_xyz_submean, 128, [1, 1], vs, 'conv-reg1', True, is_training, bn_decay=bn_decay)
        net = conv2d(net, 128, [1, 1], vs, 'conv-reg2', True, is_training, bn_decay=bn_decay)
        net = conv2d(net, 256, [1, 1], vs, 'conv-reg3', True, is_training, bn_decay=bn_decay)
        net = conv2d(net, 512, [1, 1], vs, 'conv-reg4', True, is_training, bn_decay=bn_decay)
        masked_net = net * mask
        net = max_pool_points(masked_net)
        end_points[prefix + 'feats_lv1'] = net
        if one_hot_vec is not None:
            net = torch.cat([net, one_hot_vec], dim=1)
        if norm_box2D is not None:
            net = torch.cat([net, norm_box2D], dim=1)
        net = fully_connected(net, 512, vs, 'fc1', True, is_training, bn_decay=bn_decay)
        end_points[prefix + 'feats_lv2'] = net
        net = fully_connected(net, 256, vs, 'fc2', True, is_training, bn_decay=bn_decay)
        end_points[prefix + 'feats_lv3'] = net
        output = fully_connected(net, 3 + NUM_HEADING_BIN * 2 + NUM_SIZE_CLUSTER * 4, vs, 'fc3', activation_fn=None)
        end_points[prefix + 'box_params'] = output
        return parse_box_output(output, stage1_center, end_points, prefix)


def box_pc_mask_features_model(box, pc, logits, num_outputs, is_training, end_points, reuse, bn_for_output, vs,
                               normalize_pc=False, normalize_method='SD', one_hot_vec=None, norm_box2D=None,
                               bn_decay=None, c=None, scope=None):
    """semisup_models.py:297-324."""
    if c.BOX_PC_MASK_REPRESENTATION == 'A':
        return combined_box_pc_mask_features_model(box, pc, logits, num_outputs, is_training, end_points, reuse,
                                                   False, vs, normalize_pc, normalize_method, one_hot_vec, None,
                                                   bn_decay, c, 'box_pc_mask_model')
    if c.BOX_PC_MASK_REPRESENTATION == 'B':
        return independent_box_pc_mask_features_model(box, pc, logits, num_outputs, is_training, end_points, reuse,
                                                      False, vs, normalize_pc, normalize_method, one_hot_vec, None,
                                                      bn_decay, c, 'box_pc_mask_model')
    raise Exception('Box pc mask representation not implemented: %s' % c.BOX_PC_MASK_REPRESENTATION)


def combined_box_pc_mask_features_model(box_reg, pc, mask, num_outputs, is_training, end_points, reuse,
                                        bn_for_output, vs, normalize_pc=False, normalize_method='SD',
                                        one_hot_vec=None, norm_box2D=None, bn_decay=None, c=None, scope=None):
    """semisup_models.py:326-398."""
    with vs.variable_scope(scope):
        rep = tf_util.tf_get_box_pc_representation(box_reg, pc)               # (B,N,C+6)
        if normalize_pc:
            rep = _normalize_xyz(rep, normalize_method)
        if mask is not None:
            rep = torch.cat([rep, mask], dim=2)
        D = rep.shape[2]
        net = conv2d(rep, 128, [1, D], vs, 'conv-reg1', True, is_training, bn_decay=bn_decay)
        net = conv2d(net, 128, [1, 1], vs, 'conv-reg2', True, is_training, bn_decay=bn_decay)
        net = conv2d(net, 256, [1, 1], vs, 'conv-reg3', True, is_training, bn_decay=bn_decay)
        net = conv2d(net, 512, [1, 1], vs, 'conv-reg4', True, is_training, bn_decay=bn_decay)
        if mask is not None:
            net = net * mask
        net = max_pool_points(net)
        if one_hot_vec is not None:
            net = torch.cat([net, one_hot_vec], dim=1)
        if norm_box2D is not None:
            net = torch.cat([net, norm_box2D], dim=1)
        features_lv1 = net
        net = fully_connected(net, 512, vs, 'fc1', True, is_training, bn_decay=bn_decay)
        features_lv2 = net
        net = dropout(net, vs, is_training, 'dp1', keep_prob=0.7)
        net = fully_connected(net, 256, vs, 'fc2', True, is_training, bn_decay=bn_decay)
        features_lv3 = net
        net = dropout(net, vs, is_training, 'dp2', keep_prob=0.7)
        net = fully_connected(net, num_outputs, vs, 'fc3', bn=bn_for_output, is_training=is_training,
                              activation_fn=None, bn_decay=bn_decay)
        features = {'%s_feats_lv1' % scope: features_lv1, '%s_feats_lv2' % scope: features_lv2,
                    '%s_feats_lv3' % scope: features_lv3}
    return net, features


def independent_box_pc_mask_features_model(box_reg, pc, mask, num_outputs, is_training, end_points, reuse,
                                           bn_for_output, vs, normalize_pc=False, normalize_method='SD',
                                           one_hot_vec=None, norm_box2D=None, bn_decay=None, c=None, scope=None):
    """semisup_models.py:400-471 (representation B): box parameters (B,7) through an FC stack, the raw points through a
    conv stack + max-pool, the two 512-vectors concatenated (box first; then norm_box2D, then the one-hot) into a 4-layer
    FC head."""
    with vs.variable_scope(scope):
        box7 = torch.cat([box_reg[0], box_reg[1], box_reg[2].unsqueeze(1)], dim=1)
        box_feat = mlps(box7, [128, 128, 256, 512], is_training, vs, bn=True, bn_decay=bn_decay, c=c, scope='extract_box_feats')
        if normalize_pc:
            pc = _normalize_xyz(pc, normalize_method)
        D = pc.shape[2]
        net = conv2d(pc, 128, [1, D], vs, 'conv-reg1', True, is_training, bn_decay=bn_decay)
        net = conv2d(net, 128, [1, 1], vs, 'conv-reg2', True, is_training, bn_decay=bn_decay)
        net = conv2d(net, 256, [1, 1], vs, 'conv-reg3', True, is_training, bn_decay=bn_decay)
        net = conv2d(net, 512, [1, 1], vs, 'conv-reg4', True, is_training, bn_decay=bn_decay)
        if mask is not None:
            net = net * mask
        net = max_pool_points(net)
        net = torch.cat([box_feat, net], dim=1)
        features_lv1 = net
        if norm_box2D is not None:
            net = torch.cat([net, norm_box2D], dim=1)
        if one_hot_vec is not None:
            net = torch.cat([net, one_hot_vec], dim=1)
        net = fully_connected(net, 512, vs, 'fc1', True, is_training, bn_decay=bn_decay)
        net = fully_connected(net, 512, vs, 'fc2', True, is_training, bn_decay=bn_decay)
        features_lv2 = net
        net = dropout(net, vs, is_training, 'dp2', keep_prob=0.7)
        net = fully_connected(net, 256, vs, 'fc3', True, is_training, bn_decay=bn_decay)
        features_lv3 = net
        net = dropout(net, vs, is_training, 'dp3', keep_prob=0.7)
        net = fully_connected(net, num_outputs, vs, 'fc4', bn=bn_for_output, is_training=is_training,
                              activation_fn=None, bn_decay=bn_decay)
        features = {'%s_feats_lv1' % scope: features_lv1, '%s_feats_lv2' % scope: features_lv2,
                    '%s_feats_lv3' % scope: features_lv3}
    return net, features
