"""TEST INFRASTRUCTURE (oracle): the Frustum PointNet v1 pipeline of BASELINE cfg3 composed from the restated helpers of
models/model_util.py exactly as SURVEY 3.2 lays it out -- v1_inst_seg with the one-hot (semisup_models.py:69-139),
point_cloud_masking + 512-point resample (model_util.py:241-286), get_center_regression_net (:289-325), the
box-estimation net on the gathered points (layer spec semisup_models.py:224-261, no mask multiply),
parse_output_to_tensors with NS = 10 (:178-210), center += stage1_center.  Model-A variable names (inst_seg / tnet /
box_est).  `logits` given: the pipeline continues from THOSE mask logits (used to compare everything downstream of
the strict logit compare on identical masks)."""
import torch

from . import semisup_models as osm, model_util as omu
from .tf_layers import conv2d, fully_connected, max_pool_points
from transferable3d_b200.constants import MEAN_DIMS_ARR


def get_model(vs, pc, one_hot, seed=5, logits=None, rng_mode='philox'):
    ep = {}
    if logits is None:
        logits = osm.v1_inst_seg(pc, None, one_hot, ep, False, vs, scope='inst_seg')
    obj, mean, ep = omu.point_cloud_masking(pc, logits, ep, rng_mode=rng_mode, seed=seed)
    with vs.variable_scope('tnet'):
        delta, _ = omu.get_center_regression_net(obj, one_hot, False, None, ep, vs)
    s1 = delta + mean
    with vs.variable_scope('box_est'):
        net = obj - delta.unsqueeze(1)
        for nm, c in (('conv-reg1', 128), ('conv-reg2', 128), ('conv-reg3', 256), ('conv-reg4', 512)):
            net = conv2d(net, c, [1, 1], vs, nm, True, False)
        net = torch.cat([max_pool_points(net), one_hot], dim=1)
        net = fully_connected(net, 512, vs, 'fc1', True, False)
        net = fully_connected(net, 256, vs, 'fc2', True, False)
        out = fully_connected(net, 67, vs, 'fc3', activation_fn=None)
    ep = omu.parse_output_to_tensors(out, ep, 12, MEAN_DIMS_ARR)
    ep['stage1_center'] = s1
    ep['center'] = ep['center_boxnet'] + s1
    ep['mask_logits'] = logits
    return ep
