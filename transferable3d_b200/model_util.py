"""Mirror of the F-PointNet helpers of models/model_util.py on the B200 path: tf_gather_object_pc
(:61-91), get_box3d_corners_helper / get_box3d_corners(_sunrgbd) (:94-167), parse_output_to_tensors
(:178-210, parameterised to NS / mean sizes), placeholder_inputs (:216-238), point_cloud_masking
(:241-286), get_center_regression_net (:289-325).
"""
import numpy as np
import torch

from . import runtime as rt
from . import tf_util
from ._lib import ptr, stream, call
from .constants import (NUM_HEADING_BIN, NUM_OBJECT_POINT, MEAN_DIMS_ARR, ORIENT_ANCHORS, g_mean_size_arr,
                        NUM_SIZE_CLUSTER)

# resampling RNG: 'philox' (counter-based, on device, parallel) or 'numpy_legacy' (the reference's
# sequential host numpy stream; the host draws rank-space choices from the device counts)
_rng = {'mode': 'philox', 'seed': 0, 'numpy_state': None}


def set_resample_rng(mode, seed=0):
    if mode not in ('philox', 'numpy_legacy'):
        raise ValueError(mode)
    _rng['mode'], _rng['seed'] = mode, int(seed)
    _rng['numpy_state'] = np.random.RandomState(seed) if mode == 'numpy_legacy' else None


def _numpy_legacy_choices(count, npoints):
    """The three np.random calls of model_util.py:77-84 in rank space, frustum by frustum."""
    rs = _rng['numpy_state']
    choice = np.zeros((len(count), npoints), dtype=np.int32)
    for i, n in enumerate(count):
        if n > 0:
            if n > npoints:
                ch = rs.choice(n, npoints, replace=False)
            else:
                ch = rs.choice(n, npoints - n, replace=True)
                ch = np.concatenate((np.arange(n), ch))
            rs.shuffle(ch)
            choice[i] = ch
    return choice


def _resample(pc, idx, count, mean, npoints, c_out):
    B, N, C = pc.shape
    dev = pc.device
    indices = torch.empty((B, npoints, 2), dtype=torch.int32, device=dev)
    object_pc = torch.empty((B, npoints, c_out), dtype=torch.float32, device=dev)
    if _rng['mode'] == 'philox':
        mode, choice = 0, None
    else:
        mode = 1
        choice = torch.from_numpy(_numpy_legacy_choices(count.cpu().numpy(), npoints)).to(dev)   # host round trip
    call('t3d_resample', ptr(idx), ptr(count), B, N, npoints, mode, _rng['seed'], ptr(choice), ptr(indices), ptr(pc), C,
         ptr(mean), c_out, ptr(object_pc), stream())
    return object_pc, indices


def tf_gather_object_pc(point_cloud, mask, npoints=512):
    """model_util.py:61-91: (B,N,C),(B,N) -> object_pc (B,npoints,C), indices int32 (B,npoints,2)."""
    pc = rt.f32(point_cloud)
    B, N, C = pc.shape
    m = rt.f32(mask.reshape(B, N))
    fake_logits = torch.stack([torch.full_like(m, 0.5), m], dim=2).contiguous()      # mask > 0.5
    _, count, _, _, idx = rt.mask_centroid(fake_logits, pc, want_mask=False)
    zero = torch.zeros((B, 3), dtype=torch.float32, device=pc.device)
    return _resample(pc, idx, count, zero, npoints, C)


def get_box3d_corners_helper(centers, headings, sizes):
    """model_util.py:94-119: (N,3),(N,),(N,3) -> (N,8,3)."""
    centers, headings, sizes = rt.f32(centers), rt.f32(headings), rt.f32(sizes)
    n = centers.shape[0]
    out = torch.empty((n, 8, 3), dtype=torch.float32, device=centers.device)
    call('t3d_box3d_corners_helper', ptr(centers), ptr(headings), ptr(sizes), n, ptr(out), stream())
    return out


def _corners_all(center, heading_residuals, size_residuals, mean_size_arr, nh):
    center, hr, sr = rt.f32(center), rt.f32(heading_residuals), rt.f32(size_residuals)
    B, NS = center.shape[0], mean_size_arr.shape[0]
    st = rt.store()
    ms = st.const('mean_size_%d' % NS, mean_size_arr)
    oa = st.const('orient_anchors_%d' % nh, np.arange(0, 2 * np.pi, 2 * np.pi / nh))
    out = torch.empty((B, nh, NS, 8, 3), dtype=torch.float32, device=center.device)
    call('t3d_box3d_corners_all', ptr(center), ptr(hr), ptr(sr), ptr(ms), ptr(oa), B, nh, NS, ptr(out), stream())
    return out


def get_box3d_corners(center, heading_residuals, size_residuals):
    """model_util.py:121-143 (KITTI constants; size residual added twice, as in the reference)."""
    return _corners_all(center, heading_residuals, size_residuals, g_mean_size_arr, NUM_HEADING_BIN)


def get_box3d_corners_sunrgbd(center, heading_residuals, size_residuals):
    """model_util.py:145-167."""
    return _corners_all(center, heading_residuals, size_residuals, MEAN_DIMS_ARR, NUM_HEADING_BIN)


def parse_output_to_tensors(output, end_points, num_heading_bin=NUM_HEADING_BIN, mean_size_arr=None):
    """model_util.py:178-210 (centre stored as 'center_boxnet', no stage-1 add). The module constants
    there are KITTI (NS=8); pass mean_size_arr=MEAN_DIMS_ARR for SUN-RGBD (NS=10, head width 67)."""
    if mean_size_arr is None:
        mean_size_arr = g_mean_size_arr
    st = rt.store()
    NS = mean_size_arr.shape[0]
    ms = st.const('mean_size_%d' % NS, mean_size_arr)
    oa = st.const('orient_anchors_%d' % num_heading_bin, np.arange(0, 2 * np.pi, 2 * np.pi / num_heading_bin))
    p = tf_util.parse_box_output(output, None, ms, oa, want_reg=False)
    end_points['center_boxnet'] = p['center']
    for k in ('heading_scores', 'heading_residuals_normalized', 'heading_residuals', 'size_scores',
              'size_residuals_normalized', 'size_residuals'):
        end_points[k] = p[k]
    return end_points


def placeholder_inputs(batch_size, num_point, num_channel=4, num_class=3, device='cuda'):
    """model_util.py:216-238 (KITTI widths by default: 4-channel points, 3-class one-hot)."""
    f, i = torch.float32, torch.int32
    B, N = batch_size, num_point
    Z = lambda shape, dt: torch.zeros(shape, dtype=dt, device=device)
    return (Z((B, N, num_channel), f), Z((B, num_class), f), Z((B, N), i), Z((B, 3), f), Z((B,), i), Z((B,), f),
            Z((B,), i), Z((B, 3), f))


def assemble_point_cloud(xyz, rgb_u8, out=None, lazy=False):
    """Wire format of the inference input -> the (B,N,6) fp32 point cloud of the reference placeholder
    (semisup_v1_sunrgbd.py:39): xyz (B,N,3) fp32 + rgb (B,N,3) uint8, colours = k / 255 (bit-identical to the host's
    float32 division).  15 instead of 24 bytes per point cross PCIe.  `out`: optional preallocated (B,N,6) tensor.
    lazy=True returns an rt.WirePoints instead: frustum_pointnets_v1.get_model / inference and v1_inst_seg accept it, the fused
    bf16 inst_seg chain converts the colours while it loads the points and the (B,N,6) tensor is never written."""
    xyz = rt.f32(xyz)
    if rgb_u8.dtype != torch.uint8 or tuple(rgb_u8.shape) != tuple(xyz.shape) or xyz.shape[-1] != 3:
        raise ValueError('assemble_point_cloud: xyz (B,N,3) float32 and rgb (B,N,3) uint8 expected')
    rgb_u8 = rgb_u8.contiguous()
    if lazy:
        return rt.WirePoints(xyz, rgb_u8)
    B, N = xyz.shape[0], xyz.shape[1]
    if out is None:
        out = torch.empty((B, N, 6), dtype=torch.float32, device=xyz.device)
    call('t3d_assemble_points', ptr(xyz), ptr(rgb_u8), B * N, ptr(out), stream())
    return out


def point_cloud_masking(point_cloud, logits, end_points, xyz_only=True):
    """model_util.py:241-286 -> (object_point_cloud (B,512,3|C), mask_xyz_mean (B,3), end_points)."""
    pc = rt.f32(point_cloud)
    if isinstance(pc, rt.WirePoints):
        pc = pc.xyz if xyz_only else pc.dense()          # mask, centroid and the gathered object points use xyz only
    B, N, C = pc.shape
    mask, count, mean, _, idx = rt.mask_centroid(logits, pc)
    end_points['mask'] = mask
    c_out = 3 if xyz_only else C
    object_pc, indices = _resample(pc, idx, count, mean, NUM_OBJECT_POINT, c_out)
    end_points['object_pc_indices'] = indices
    return object_pc, mean, end_points


def get_center_regression_net(object_point_cloud, one_hot_vec, is_training, bn_decay, end_points):
    """model_util.py:289-325: T-Net on the gathered object points (variables in the current scope)."""
    rt.require_eval(is_training)
    from .semisup_models import PointView, _masked_chain, _cat
    st = rt.store()
    full = st.scope_name()
    net = _masked_chain(rt.CHAIN_TNET, PointView(rt.f32(object_point_cloud), None), None,
                        ['conv-reg1-stage1', 'conv-reg2-stage1', 'conv-reg3-stage1'], full)
    net = _cat([net, one_hot_vec])
    net = tf_util.fully_connected(net, 256, scope='fc1-stage1', bn=True, is_training=is_training, bn_decay=bn_decay)
    net = tf_util.fully_connected(net, 128, scope='fc2-stage1', bn=True, is_training=is_training, bn_decay=bn_decay)
    predicted_center = tf_util.fully_connected(net, 3, activation_fn=None, scope='fc3-stage1')
    return predicted_center, end_points
