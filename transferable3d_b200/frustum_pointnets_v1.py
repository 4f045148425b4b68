"""Frustum PointNet v1 pipeline (BASELINE cfg3), composed from the inherited helpers of
models/model_util.py exactly as SURVEY 3.2 lays out (the reference repo carries the helpers but no
driver for them): v1_inst_seg (one-hot in the global feature) -> point_cloud_masking (mask, centroid,
512-point resample, model_util.py:241-286) -> get_center_regression_net (:289-325) -> box-estimation
net on the 512 object points (layer spec semisup_models.py:224-261 without the mask multiply) ->
parse_output_to_tensors with NS=10 / SUN mean sizes (:178-210), center += stage1_center.
Variable scopes: inst_seg, tnet, box_est (model-A names).
"""
import torch

from . import runtime as rt
from . import tf_util, semisup_models, model_util
from .semisup_models import PointView, _masked_chain, _cat
from .constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER, MEAN_DIMS_ARR


def get_model(point_cloud, one_hot_vec, is_training, bn_decay=None, end_points=None):
    rt.require_eval(is_training)
    end_points = {} if end_points is None else end_points
    pc = rt.f32(point_cloud)
    logits = semisup_models.v1_inst_seg(pc, None, one_hot_vec, end_points, is_training, bn_decay=bn_decay, scope='inst_seg')
    end_points['mask_logits'] = logits
    object_pc, mask_xyz_mean, end_points = model_util.point_cloud_masking(pc, logits, end_points)
    with rt.variable_scope('tnet'):
        center_delta, end_points = model_util.get_center_regression_net(object_pc, one_hot_vec, is_training, bn_decay,
                                                                         end_points)
    stage1_center = center_delta + mask_xyz_mean
    end_points['stage1_center'] = stage1_center
    with rt.variable_scope('box_est'):
        full = rt.store().scope_name()
        # object_pc - center_delta is applied on load by the fused kernel
        net = _masked_chain(rt.CHAIN_BOX, PointView(object_pc, center_delta), None,
                            ['conv-reg1', 'conv-reg2', 'conv-reg3', 'conv-reg4'], full)
        net = _cat([net, one_hot_vec])
        net = tf_util.fully_connected(net, 512, scope='fc1', bn=True, is_training=is_training, bn_decay=bn_decay)
        net = tf_util.fully_connected(net, 256, scope='fc2', bn=True, is_training=is_training, bn_decay=bn_decay)
        output = tf_util.fully_connected(net, 3 + NUM_HEADING_BIN * 2 + NUM_SIZE_CLUSTER * 4, activation_fn=None, scope='fc3')
    end_points = model_util.parse_output_to_tensors(output, end_points, NUM_HEADING_BIN, MEAN_DIMS_ARR)
    end_points['center'] = end_points['center_boxnet'] + stage1_center
    return end_points


def inference(point_cloud, one_hot_vec, end_points=None):
    """get_model + the reference's test-time post-processing (test_semisup.inference, test_semisup.py:236-258: softmax over
    the point logits, mask_mean_prob, log-score, argmax-selected residuals) on the device.  Returns the prediction the
    reference's runner returns -- a dict of device tensors pred_seg (B,N) uint8, center (B,3), heading_cls, heading_res,
    size_cls, size_res (B,3), scores -- i.e. N bytes + 10 numbers per frustum instead of the 2 N + 67 raw fetches."""
    from .test_semisup import inference_scores
    ep = get_model(point_cloud, one_hot_vec, False, end_points=end_points)
    out = inference_scores(ep['mask_logits'], ep['heading_scores'], ep['heading_residuals'], ep['size_scores'], ep['size_residuals'])
    out['center'] = ep['center']
    out['end_points'] = ep
    return out
