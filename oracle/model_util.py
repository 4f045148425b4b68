"""F-PointNet shared helpers: restatement of models/model_util.py:61-448.

Resampling RNG (model_util.py:71-87) is an explicit input with two modes:
  * 'numpy_legacy': literal -- consumes a numpy RandomState frustum by frustum exactly like the
    reference's global np.random stream (choice / choice / shuffle).
  * 'philox': counter-based restatement of the same procedure (random ordered subset when
    count > npoints; identity ++ uniform refill, then a random shuffle otherwise) with
    Philox4x32-10 keyed per (seed, frustum, element) so a GPU can reproduce it bit-exactly in
    parallel.  Selection/shuffle = stable sort by 64-bit Philox key (ties by position).
"""
import numpy as np
import torch

from . import tf_layers
from .tf_layers import conv2d, fully_connected, max_pool_points

NUM_HEADING_BIN = 12
NUM_OBJECT_POINT = 512

# ----------------------------------------------------------------------------- Philox4x32-10

_PH_M0, _PH_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PH_W0, _PH_W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Standard Philox4x32-10 (Salmon et al. 2011) on uint32 numpy arrays; returns 4 words."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = _PH_M0 * c0.astype(np.uint64)
            p1 = _PH_M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_PH_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_PH_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _philox_key64(seed, frustum, stream, pos):
    w0, w1, _, _ = philox4x32_10(pos, np.uint32(stream), np.uint32(frustum), np.uint32(0),
                                 seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return (w0.astype(np.uint64) << np.uint64(32)) | w1.astype(np.uint64)


def philox_choice(count, npoints, seed, frustum):
    """Rank-space `choice` array (positions into pos_indices) for one frustum, count > 0."""
    if count > npoints:
        keys = _philox_key64(seed, frustum, 0, np.arange(count, dtype=np.uint32))
        order = np.argsort(keys, kind='stable')
        return order[:npoints].astype(np.int64)
    t = np.arange(npoints, dtype=np.uint32)
    _, _, w2, _ = philox4x32_10(t, np.uint32(1), np.uint32(frustum), np.uint32(0),
                                seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    lst = np.where(t < count, t, w2 % np.uint32(count)).astype(np.int64)
    keys = _philox_key64(seed, frustum, 2, t)
    order = np.argsort(keys, kind='stable')
    return lst[order]


def mask_to_indices(mask, npoints=NUM_OBJECT_POINT, rng_mode='numpy_legacy', rng=None, seed=0):
    """model_util.py:71-87. mask: (B,N) numpy of 0/1. Returns int32 (B,npoints,2)."""
    mask = np.asarray(mask)
    indices = np.zeros((mask.shape[0], npoints, 2), dtype=np.int32)
    for i in range(mask.shape[0]):
        pos_indices = np.where(mask[i, :] > 0.5)[0]
        if len(pos_indices) > 0:
            if rng_mode == 'numpy_legacy':
                if len(pos_indices) > npoints:
                    choice = rng.choice(len(pos_indices), npoints, replace=False)
                else:
                    choice = rng.choice(len(pos_indices), npoints - len(pos_indices), replace=True)
                    choice = np.concatenate((np.arange(len(pos_indices)), choice))
                rng.shuffle(choice)
            elif rng_mode == 'philox':
                choice = philox_choice(len(pos_indices), npoints, seed, i)
            elif rng_mode == 'choice':      # caller-supplied rank-space choice arrays
                choice = np.asarray(rng[i])
            else:
                raise ValueError(rng_mode)
            indices[i, :, 1] = pos_indices[choice]
        indices[i, :, 0] = i
    return indices


def tf_gather_object_pc(point_cloud, mask, npoints=512, **rng_kw):
    """model_util.py:61-91: (B,N,C),(B,N) -> (B,npoints,C), int32 (B,npoints,2)."""
    indices = mask_to_indices(mask.detach().cpu().numpy(), npoints, **rng_kw)
    idx = torch.as_tensor(indices.astype(np.int64))
    object_pc = point_cloud[idx[:, :, 0], idx[:, :, 1]]
    return object_pc, indices


# ----------------------------------------------------------------------------- boxes

def get_box3d_corners_helper(centers, headings, sizes):
    """model_util.py:94-119: (N,3),(N,),(N,3) -> (N,8,3)."""
    l, w, h = sizes[:, 0:1], sizes[:, 1:2], sizes[:, 2:3]
    x_c = torch.cat([l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2, -l / 2], dim=1)
    y_c = torch.cat([h / 2, h / 2, h / 2, h / 2, -h / 2, -h / 2, -h / 2, -h / 2], dim=1)
    z_c = torch.cat([w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2, w / 2], dim=1)
    corners = torch.stack([x_c, y_c, z_c], dim=1)                               # (N,3,8)
    c, s = torch.cos(headings), torch.sin(headings)
    ones, zeros = torch.ones_like(c), torch.zeros_like(c)
    R = torch.stack([torch.stack([c, zeros, s], dim=1), torch.stack([zeros, ones, zeros], dim=1),
                     torch.stack([-s, zeros, c], dim=1)], dim=1)
    corners_3d = torch.matmul(R, corners) + centers.unsqueeze(2)
    return corners_3d.transpose(1, 2)


def get_box3d_corners(center, heading_residuals, size_residuals, mean_size_arr, num_heading_bin=NUM_HEADING_BIN):
    """model_util.py:121-143 / :145-167 (SUN-RGBD variant = same code with the SUN constants).
    NOTE the reference adds the size residual twice (:134-135 / :158-159); replicated."""
    B = center.shape[0]
    NH = num_heading_bin
    NS = mean_size_arr.shape[0]
    dt = center.dtype
    bins = torch.as_tensor(np.arange(0, 2 * np.pi, 2 * np.pi / NH), dtype=torch.float32).to(dt)
    headings = heading_residuals + bins.unsqueeze(0)                            # (B,NH)
    mean_sizes = torch.as_tensor(mean_size_arr, dtype=torch.float32).to(dt).unsqueeze(0) + size_residuals
    sizes = mean_sizes + size_residuals                                         # (B,NS,3)
    sizes = sizes.unsqueeze(1).expand(B, NH, NS, 3)
    headings = headings.unsqueeze(-1).expand(B, NH, NS)
    centers = center.unsqueeze(1).unsqueeze(1).expand(B, NH, NS, 3)
    n = B * NH * NS
    c3 = get_box3d_corners_helper(centers.reshape(n, 3), headings.reshape(n), sizes.reshape(n, 3))
    return c3.reshape(B, NH, NS, 8, 3)


def get_box3d_corners_sunrgbd(center, heading_residuals, size_residuals):
    from transferable3d_b200.constants import MEAN_DIMS_ARR
    return get_box3d_corners(center, heading_residuals, size_residuals, MEAN_DIMS_ARR, 12)


def huber_loss(error, delta):
    """model_util.py:170-175 (mean-reduced)."""
    abs_error = error.abs()
    quadratic = torch.clamp(abs_error, max=delta)
    linear = abs_error - quadratic
    return (0.5 * quadratic ** 2 + delta * linear).mean()


def parse_output_to_tensors(output, end_points, num_heading_bin=NUM_HEADING_BIN, mean_size_arr=None):
    """model_util.py:178-210, parameterised over (NH, mean size table): the module constants
    there are KITTI (NS=8); this task uses NS=10 + sun_mean_size_arr (model_util.py:37-54)."""
    if mean_size_arr is None:
        from transferable3d_b200.constants import g_mean_size_arr as mean_size_arr
    NH, NS = num_heading_bin, mean_size_arr.shape[0]
    B = output.shape[0]
    end_points['center_boxnet'] = output[:, 0:3]
    end_points['heading_scores'] = output[:, 3:3 + NH]
    hrn = output[:, 3 + NH:3 + 2 * NH]
    end_points['heading_residuals_normalized'] = hrn
    end_points['heading_residuals'] = hrn * (np.pi / NH)
    end_points['size_scores'] = output[:, 3 + 2 * NH:3 + 2 * NH + NS]
    srn = output[:, 3 + 2 * NH + NS:3 + 2 * NH + 4 * NS].reshape(B, NS, 3)
    end_points['size_residuals_normalized'] = srn
    end_points['size_residuals'] = srn * torch.as_tensor(mean_size_arr, dtype=torch.float32).to(output.dtype).unsqueeze(0)
    return end_points


# ----------------------------------------------------------------------------- shared subgraphs

def point_cloud_masking(point_cloud, logits, end_points, xyz_only=True, npoints=NUM_OBJECT_POINT, **rng_kw):
    """model_util.py:241-286."""
    mask = (logits[:, :, 0:1] < logits[:, :, 1:2]).to(point_cloud.dtype)       # (B,N,1)
    mask_count = mask.sum(dim=1, keepdim=True).repeat(1, 1, 3)
    xyz = point_cloud[:, :, 0:3]
    mean = (mask.repeat(1, 1, 3) * xyz).sum(dim=1, keepdim=True)
    mask2 = mask.squeeze(2)
    end_points['mask'] = mask2
    mean = mean / torch.clamp(mask_count, min=1)
    xyz_stage1 = xyz - mean
    if xyz_only:
        stage1 = xyz_stage1
    else:
        stage1 = torch.cat([xyz_stage1, point_cloud[:, :, 3:]], dim=-1)
    object_pc, indices = tf_gather_object_pc(stage1, mask2, npoints, **rng_kw)
    end_points['object_pc_indices'] = indices       # not in the reference dict; test hook
    return object_pc, mean.squeeze(1), end_points


def get_center_regression_net(object_point_cloud, one_hot_vec, is_training, bn_decay, end_points, vs):
    """model_util.py:289-325 (T-Net on the gathered object points)."""
    net = conv2d(object_point_cloud, 128, [1, 1], vs, 'conv-reg1-stage1', True, is_training, bn_decay=bn_decay)
    net = conv2d(net, 128, [1, 1], vs, 'conv-reg2-stage1', True, is_training, bn_decay=bn_decay)
    net = conv2d(net, 256, [1, 1], vs, 'conv-reg3-stage1', True, is_training, bn_decay=bn_decay)
    net = max_pool_points(net)
    if one_hot_vec is not None:
        net = torch.cat([net, one_hot_vec], dim=1)
    net = fully_connected(net, 256, vs, 'fc1-stage1', True, is_training, bn_decay=bn_decay)
    net = fully_connected(net, 128, vs, 'fc2-stage1', True, is_training, bn_decay=bn_decay)
    predicted_center = fully_connected(net, 3, vs, 'fc3-stage1', activation_fn=None)
    return predicted_center, end_points


def get_loss(mask_label, center_label, heading_class_label, heading_residual_label,
             size_class_label, size_residual_label, end_points,
             corner_loss_weight=10.0, box_loss_weight=1.0, mean_size_arr=None,
             num_heading_bin=NUM_HEADING_BIN):
    """model_util.py:328-448 (mean-reduced strong loss of F-PointNet v1)."""
    if mean_size_arr is None:
        from transferable3d_b200.constants import g_mean_size_arr as mean_size_arr
    F = torch.nn.functional
    NH, NS = num_heading_bin, mean_size_arr.shape[0]
    dt = end_points['center'].dtype
    msa = torch.as_tensor(mean_size_arr, dtype=torch.float32).to(dt)
    logits = end_points['mask_logits']
    mask_loss = F.cross_entropy(logits.reshape(-1, 2), mask_label.reshape(-1).long())
    center_dist = torch.linalg.norm(center_label - end_points['center'], dim=-1)
    center_loss = huber_loss(center_dist, 2.0)
    s1_dist = torch.linalg.norm(center_label - end_points['stage1_center'], dim=-1)
    stage1_center_loss = huber_loss(s1_dist, 1.0)
    heading_class_loss = F.cross_entropy(end_points['heading_scores'], heading_class_label.long())
    hcls = F.one_hot(heading_class_label.long(), NH).to(dt)
    hrn_label = heading_residual_label / (np.pi / NH)
    hrn_loss = huber_loss((end_points['heading_residuals_normalized'] * hcls).sum(dim=1) - hrn_label, 1.0)
    size_class_loss = F.cross_entropy(end_points['size_scores'], size_class_label.long())
    scls = F.one_hot(size_class_label.long(), NS).to(dt)
    scls_t = scls.unsqueeze(-1).repeat(1, 1, 3)
    pred_srn = (end_points['size_residuals_normalized'] * scls_t).sum(dim=1)
    mean_size_label = (scls_t * msa.unsqueeze(0)).sum(dim=1)
    srl_norm = size_residual_label / mean_size_label
    srn_loss = huber_loss(torch.linalg.norm(srl_norm - pred_srn, dim=-1), 1.0)
    corners_3d = get_box3d_corners(end_points['center'], end_points['heading_residuals'],
                                   end_points['size_residuals'], mean_size_arr, NH)
    gt_mask = hcls.unsqueeze(2) * scls.unsqueeze(1)
    corners_pred = (gt_mask.unsqueeze(-1).unsqueeze(-1) * corners_3d).sum(dim=(1, 2))
    bins = torch.as_tensor(np.arange(0, 2 * np.pi, 2 * np.pi / NH), dtype=torch.float32).to(dt)
    heading_label = (hcls * (heading_residual_label.unsqueeze(1) + bins.unsqueeze(0))).sum(dim=1)
    size_label = (scls.unsqueeze(-1) * (msa.unsqueeze(0) + size_residual_label.unsqueeze(1))).sum(dim=1)
    c_gt = get_box3d_corners_helper(center_label, heading_label, size_label)
    c_gt_flip = get_box3d_corners_helper(center_label, heading_label + np.pi, size_label)
    corners_dist = torch.minimum(torch.linalg.norm(corners_pred - c_gt, dim=-1),
                                 torch.linalg.norm(corners_pred - c_gt_flip, dim=-1))
    corners_loss = huber_loss(corners_dist, 1.0)
    return mask_loss + box_loss_weight * (center_loss + heading_class_loss + size_class_loss +
                                          hrn_loss * 20 + srn_loss * 20 + stage1_center_loss +
                                          corner_loss_weight * corners_loss)
