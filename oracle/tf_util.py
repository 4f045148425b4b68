"""Geometry / box ops: restatement of models/tf_util.py:364-540 (2D box ops), :764-1073
(BoxPC representation, SUN-RGBD projection, box construction, anchor->reg, frustum rotation)
and :1141 (tf_expand_tile).  tf.map_fn over the batch is replaced by batched tensor ops with
identical per-sample arithmetic.
"""
import torch


def tf_expand_tile(tensor, axis, tile):
    """tf_util.py:1141-1142."""
    return tensor.unsqueeze(axis).repeat(*tile)


# ----------------------------------------------------------------------------- 2D boxes

def tf_get_2D_bbox_of_points(points2D):
    """tf_util.py:364-377. (..., P, 2) -> (..., 4) = [left, top, right, bottom]."""
    left = points2D[..., 0].min(dim=-1).values
    top = points2D[..., 1].min(dim=-1).values
    right = points2D[..., 0].max(dim=-1).values
    bottom = points2D[..., 1].max(dim=-1).values
    return torch.stack([left, top, right, bottom], dim=-1)


def tf_get_2D_softmax_bbox_of_points(points2D, softmax_scale_factor):
    """tf_util.py:379-413; stop_gradient on width/height and on the softmax weights."""
    x, y = points2D[..., 0], points2D[..., 1]
    lb, rb = x.min(dim=-1, keepdim=True).values, x.max(dim=-1, keepdim=True).values
    tb, bb = y.min(dim=-1, keepdim=True).values, y.max(dim=-1, keepdim=True).values
    width = (rb - lb).abs().detach()
    height = (bb - tb).abs().detach()
    sm = lambda t: torch.softmax(t * softmax_scale_factor, dim=-1).detach()
    left = (x * sm((rb - x) / width)).sum(dim=-1)
    top = (y * sm((bb - y) / height)).sum(dim=-1)
    right = (x * sm((x - lb) / width)).sum(dim=-1)
    bottom = (y * sm((y - tb) / height)).sum(dim=-1)
    return torch.stack([left, top, right, bottom], dim=-1)


def tf_normalize_2D_bboxes(box2D, image_dim):
    """tf_util.py:466-484. image_dim = (rows, cols)."""
    rows, cols = image_dim[:, 0], image_dim[:, 1]
    return torch.stack([box2D[:, 0] / cols, box2D[:, 1] / rows,
                        box2D[:, 2] / cols, box2D[:, 3] / rows], dim=1)


def tf_dilate_2D_bboxes(bbox2D, dilate_factor):
    """tf_util.py:486-514 (height = top - bottom is negative in image coords; kept)."""
    left, top, right, bottom = bbox2D[:, 0], bbox2D[:, 1], bbox2D[:, 2], bbox2D[:, 3]
    cx, cy = (left + right) / 2., (top + bottom) / 2.
    new_w = dilate_factor * (right - left)
    new_h = dilate_factor * (top - bottom)
    return torch.stack([cx - new_w / 2., cy + new_h / 2., cx + new_w / 2., cy - new_h / 2.], dim=1)


def tf_clip_2D_bbox_to_image_dims_multi(box2Ds, image_dims):
    """tf_util.py:517-540. clip to [0,cols]x[0,rows]."""
    rows, cols = image_dims[:, 0], image_dims[:, 1]
    zero = torch.zeros_like(rows)
    return torch.stack([torch.maximum(zero, box2Ds[:, 0]), torch.maximum(zero, box2Ds[:, 1]),
                        torch.minimum(cols, box2Ds[:, 2]), torch.minimum(rows, box2Ds[:, 3])], dim=1)


# ----------------------------------------------------------------------------- SUN-RGBD projection

def flip_axis_to_camera(pc):
    """tf_util.py:816-823: depth (x,y,z) -> camera (x,-z,y)."""
    return torch.stack([pc[..., 0], -pc[..., 2], pc[..., 1]], dim=-1)


def flip_axis_to_depth(pc):
    """tf_util.py:826-830: camera (x,y,z) -> depth (x,z,-y)."""
    return torch.stack([pc[..., 0], pc[..., 2], -pc[..., 1]], dim=-1)


project_upright_depth_to_upright_camera = flip_axis_to_camera      # tf_util.py:833-834
project_upright_camera_to_upright_depth = flip_axis_to_depth       # tf_util.py:837-838


def project_upright_depth_to_camera(pc, Rtilt):
    """tf_util.py:798-804: Rtilt^T . pc, then flip to camera axes."""
    pc2 = torch.matmul(Rtilt.transpose(1, 2), pc[:, :, 0:3].transpose(1, 2))   # (B,3,N)
    return flip_axis_to_camera(pc2.transpose(1, 2))


def project_upright_depth_to_image(pc, Rtilt, K):
    """tf_util.py:807-813."""
    pc2 = project_upright_depth_to_camera(pc, Rtilt)
    uv = torch.matmul(pc2, K.transpose(1, 2))
    uv = torch.stack([uv[:, :, 0] / uv[:, :, 2], uv[:, :, 1] / uv[:, :, 2]], dim=2)
    return uv, pc2[:, :, 2]


def tf_get_2D_bbox_of_projection_sunrgbd_multi(point_clouds, Rtilts, Ks):
    """tf_util.py:416-431."""
    pts = project_upright_camera_to_upright_depth(point_clouds)
    proj, _ = project_upright_depth_to_image(pts, Rtilts, Ks)
    return tf_get_2D_bbox_of_points(proj)


def tf_get_2D_bbox_of_softmax_projection_sunrgbd_multi(point_clouds, Rtilts, Ks, softmax_scale_factor):
    """tf_util.py:434-449."""
    pts = project_upright_camera_to_upright_depth(point_clouds)
    proj, _ = project_upright_depth_to_image(pts, Rtilts, Ks)
    return tf_get_2D_softmax_bbox_of_points(proj, softmax_scale_factor)


# ----------------------------------------------------------------------------- 3D boxes

def tf_create_3D_box_by_vertices_multi(box_params, apply_translation=False):
    """tf_util.py:841-891: corners in upright depth with rotz(-theta), flipped to upright camera."""
    centers, dims_reg, orient_reg = box_params
    l, w, h = dims_reg[:, 0:1], dims_reg[:, 1:2], dims_reg[:, 2:3]
    c, s = torch.cos(-1 * orient_reg), torch.sin(-1 * orient_reg)
    x_c = torch.cat([-l / 2, l / 2, l / 2, -l / 2, -l / 2, l / 2, l / 2, -l / 2], dim=1)
    y_c = torch.cat([w / 2, w / 2, -w / 2, -w / 2, w / 2, w / 2, -w / 2, -w / 2], dim=1)
    z_c = torch.cat([h / 2, h / 2, h / 2, h / 2, -h / 2, -h / 2, -h / 2, -h / 2], dim=1)
    corners = torch.stack([x_c, y_c, z_c], dim=1)                              # (N,3,8)
    zeros, ones = torch.zeros_like(c), torch.ones_like(c)
    R = torch.stack([torch.stack([c, -s, zeros], dim=1), torch.stack([s, c, zeros], dim=1),
                     torch.stack([zeros, zeros, ones], dim=1)], dim=1)         # (N,3,3)
    corners_3d = torch.matmul(R, corners).transpose(1, 2)                      # (N,8,3)
    corners_3d = project_upright_depth_to_upright_camera(corners_3d)
    if apply_translation:
        corners_3d = corners_3d + centers.unsqueeze(1)
    return centers, corners_3d


def tf_create_3D_box_by_surface_centers_multi(box_params, apply_translation=False):
    """tf_util.py:893-955: (centers (B,3), surface points (B,6,3), inward normals (B,6,3))."""
    center, dims_reg, orient_reg = box_params
    l, w, h = dims_reg[:, 0], dims_reg[:, 1], dims_reg[:, 2]
    st, ct = torch.sin(orient_reg), torch.cos(orient_reg)
    z, o = torch.zeros_like(st), torch.ones_like(st)
    rot = torch.stack([torch.stack([ct, z, st], dim=1), torch.stack([z, o, z], dim=1),
                       torch.stack([-st, z, ct], dim=1)], dim=1)               # (B,3,3)
    sp = torch.stack([torch.stack([l / 2, -l / 2, z, z, z, z], dim=1),
                      torch.stack([z, z, h / 2, -h / 2, z, z], dim=1),
                      torch.stack([z, z, z, z, w / 2, -w / 2], dim=1)], dim=1)  # (B,3,6)
    rsp = torch.matmul(rot, sp).transpose(1, 2)                                # (B,6,3)
    sn = torch.tensor([[-1., 1., 0., 0., 0., 0.], [0., 0., -1., 1., 0., 0.],
                       [0., 0., 0., 0., -1., 1.]], dtype=rot.dtype)
    rsn = torch.matmul(rot, sn.unsqueeze(0).expand(rot.shape[0], 3, 6)).transpose(1, 2)
    if apply_translation:
        rsp = rsp + center.unsqueeze(1)
    return center, rsp, rsn


def tf_distance_to_box_surfaces_multi(point_clouds, box_params):
    """tf_util.py:610-677 batched over frustums (the reference maps the single-box function over the batch, :711-720):
    (B,N,3), (center (B,3), dims (B,3), orient (B,)) -> dist_points_to_surfaces (B,N,6).  Literal restatement: surface
    points / normals from tf_create_3D_box_by_surface_centers(apply_translation=True, use_base=False)."""
    center, surface_pts, surface_norms = tf_create_3D_box_by_surface_centers_multi(box_params, apply_translation=True)
    ray = point_clouds - center.unsqueeze(1)                                    # (B,N,3)
    perp = torch.einsum('bnc,bsc->bns', ray, surface_norms)                     # l . n
    p0l0n = ((surface_pts - center.unsqueeze(1)) * surface_norms).sum(dim=2)    # (B,6)
    norm = torch.sqrt((ray * ray).sum(dim=2))                                   # (B,N)
    dcs = p0l0n.unsqueeze(1) / (perp + 1e-5)
    dcs = norm.unsqueeze(2) * dcs
    l, w, h = box_params[1][:, 0], box_params[1][:, 1], box_params[1][:, 2]
    half = torch.stack([l / 2, l / 2, h / 2, h / 2, w / 2, w / 2], dim=1)       # (B,6)
    at_centre = (ray.abs().sum(dim=2) == 0).unsqueeze(2)
    dcs = torch.where(at_centre, half.unsqueeze(1).expand_as(dcs), dcs)
    return (norm.unsqueeze(2) - dcs).abs()


def tf_distance_to_closest_3D_box_surface_multi(point_clouds, box_params):
    """tf_util.py:679-720: the reference cleans the distances (NaN / intersection outside the box -> 1e8) and then takes
    the minimum of the UNCLEANED tensor (:707) -- replicated."""
    return tf_distance_to_box_surfaces_multi(point_clouds, box_params).min(dim=2).values


def tf_normalize_point_clouds_to_01(point_clouds):
    """tf_util.py:134-155: subtract the per-cloud centroid, divide by the largest xyz extent (+1e-5); channels >= 3 are kept."""
    centroids = point_clouds.mean(dim=1, keepdim=True)
    translated = point_clouds - centroids
    max_dims = translated.max(dim=1).values - translated.min(dim=1).values           # (B,D)
    max_dims = max_dims[:, :3].max(dim=1).values.reshape(-1, 1, 1)
    normalized = translated / (max_dims + 1e-5)
    return torch.cat([normalized[:, :, :3], point_clouds[:, :, 3:]], dim=2)


def tf_normalize_point_clouds_to_mean_zero_and_unit_var(point_clouds):
    """tf_util.py:157-173: (x - mean) / (sqrt(var) + 1e-5) per cloud and channel (tf.nn.moments: biased variance); channels
    >= 3 are kept."""
    centroids = point_clouds.mean(dim=1, keepdim=True)
    variances = ((point_clouds - centroids) ** 2).mean(dim=1, keepdim=True)
    normalized = (point_clouds - centroids) / (variances.sqrt() + 1e-5)
    return torch.cat([normalized[:, :, :3], point_clouds[:, :, 3:]], dim=2)


def tf_get_box_pc_representation(box_reg, pc):
    """tf_util.py:764-795: original pc (all channels, untranslated) ++ 6 signed plane distances."""
    center, dims_reg, orient_reg = box_reg
    translated = pc[:, :, 0:3] - center.unsqueeze(1)                            # (B,N,3)
    _, surface_pts, surface_norms = tf_create_3D_box_by_surface_centers_multi(box_reg, False)
    ray = translated.unsqueeze(2) - surface_pts.unsqueeze(1)                    # (B,N,6,3)
    perp = (surface_norms.unsqueeze(1) * ray).sum(dim=3)                        # (B,N,6)
    return torch.cat([pc, perp], dim=2)


def tf_convert_box_params_from_anchor_to_reg_format_multi(box_params, y_classes, dims_anchors, orient_anchors):
    """tf_util.py:1001-1041. argmax = first max; dims clamped at 1e-5; y_classes unused (as in
    the reference)."""
    center, dims_cls, dims_reg, orient_cls, orient_reg = box_params
    B = center.shape[0]
    ar = torch.arange(B)
    i = torch.argmax(dims_cls, dim=1)
    dims = dims_anchors.to(dims_reg.dtype)[i] + dims_reg[ar, i]
    dims = torch.clamp(dims, min=1e-5)
    j = torch.argmax(orient_cls, dim=1)
    orient = orient_anchors.to(orient_reg.dtype)[j] + orient_reg[ar, j]
    return center, dims, orient


def tf_rot_box_params_multi(box_params, angles):
    """tf_util.py:1045-1073: rotate centre about y by `angle`, orient += angle. angles: (B,1)."""
    center, dims, orient = box_params
    a = angles.reshape(-1)
    ct, st = torch.cos(a), torch.sin(a)
    x, y, z = center[:, 0], center[:, 1], center[:, 2]
    new_center = torch.stack([ct * x + st * z, y, -st * x + ct * z], dim=1)
    return new_center, dims, orient + a
