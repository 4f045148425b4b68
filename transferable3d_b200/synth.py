"""Seeded synthetic SUN-RGBD-shaped frustum batches (there is no dataset in this environment).

Shapes and conventions mirror what the reference feeds its placeholders
(sunrgbd_detection/semisup_v1_sunrgbd.py:37-67, train_semisup_adv.py:588-603):
  * a frustum is N points x C=6 channels (xyz in the centred upright-camera frame: x right,
    y down, z forward; rgb in [0,1]) -- sunrgbd_data.py:76-128, roi_seg_box3d_dataset.py:346-368
  * box labels use the class/residual codecs of roi_seg_box3d_dataset.py:47-82
  * rot_frust = pi/2 + frustum_angle (roi_seg_box3d_dataset.py:346-347), frustum_angle =
    -atan2(z,x) of the 2D-box centre ray, so rot_frust is a small angle around 0
  * Rtilt / K / img_dim=(rows, cols) follow sunrgbd_data/utils.py:38-129
All randomness comes from numpy.random.Generator(PCG64(seed)); everything is float32/int32.
"""
import numpy as np

from .constants import (NUM_HEADING_BIN, NUM_SIZE_CLUSTER, NUM_CLASS, MEAN_DIMS_ARR)


def angle2class(angle, num_class=NUM_HEADING_BIN):
    """Vectorised roi_seg_box3d_dataset.py:47-62."""
    angle = np.mod(angle, 2 * np.pi)
    per = 2 * np.pi / float(num_class)
    shifted = np.mod(angle + per / 2, 2 * np.pi)
    cls = (shifted / per).astype(np.int32)
    res = shifted - (cls * per + per / 2)
    return cls, res


def class2angle(cls, residual, num_class=NUM_HEADING_BIN, to_label_format=True):
    """Vectorised roi_seg_box3d_dataset.py:64-71."""
    per = 2 * np.pi / float(num_class)
    angle = cls * per + residual
    if to_label_format:
        angle = np.where(angle > np.pi, angle - 2 * np.pi, angle)
    return angle


def _box_corners_upright_camera(center, size, heading):
    """get_3d_box (roi_seg_box3d_dataset.py:84-100), batched: (B,3),(B,3),(B,) -> (B,8,3)."""
    l, w, h = size[:, 0:1], size[:, 1:2], size[:, 2:3]
    sx = np.array([1, 1, -1, -1, 1, 1, -1, -1], np.float64)[None] * l / 2
    sy = np.array([1, 1, 1, 1, -1, -1, -1, -1], np.float64)[None] * h / 2
    sz = np.array([1, -1, -1, 1, 1, -1, -1, 1], np.float64)[None] * w / 2
    c, s = np.cos(heading)[:, None], np.sin(heading)[:, None]
    x = c * sx + s * sz + center[:, 0:1]
    y = sy + center[:, 1:2]
    z = -s * sx + c * sz + center[:, 2:3]
    return np.stack([x, y, z], axis=2)


def project_upright_camera_to_image(pts, Rtilt, K):
    """(B,P,3) upright-camera points -> (B,P,2) pixels; chain of models/tf_util.py:798-838:
    flip to upright depth (x,z,-y), Rtilt^T, flip to camera (x,-z,y), K, divide."""
    d = np.stack([pts[..., 0], pts[..., 2], -pts[..., 1]], axis=-1)
    q = np.einsum('bji,bpj->bpi', Rtilt, d)            # Rtilt^T d
    c = np.stack([q[..., 0], -q[..., 2], q[..., 1]], axis=-1)
    uvw = np.einsum('bij,bpj->bpi', K, c)
    return uvw[..., :2] / uvw[..., 2:3]


def make_batch(B, N=2048, C=6, seed=1234, is_data_2D=0):
    """Returns a dict keyed like the placeholder tuple of semisup_v1_sunrgbd.placeholder_inputs
    (unused bg_pc/img/R0_rect/P are omitted)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cls = rng.integers(0, NUM_CLASS, size=B)
    size = MEAN_DIMS_ARR[cls] * rng.uniform(0.8, 1.2, size=(B, 3))          # (l,w,h)
    heading = rng.uniform(-np.pi, np.pi, size=B)
    center = np.stack([rng.normal(0, 0.15, size=B), rng.uniform(-0.3, 0.8, size=B),
                       rng.uniform(1.5, 6.0, size=B)], axis=1)

    # object / background split, at least 5 object points (sunrgbd_data.py:169)
    frac = rng.uniform(0.15, 0.7, size=(B, 1))
    seg = rng.random((B, N)) < frac
    seg[:, :5] = True

    # object points: on the box surface + N(0, 0.01) noise
    loc = rng.uniform(-0.5, 0.5, size=(B, N, 3))
    axis = rng.integers(0, 3, size=(B, N))
    sign = rng.integers(0, 2, size=(B, N)) * 1.0 - 0.5
    np.put_along_axis(loc, axis[..., None], sign[..., None], axis=2)
    lhw = np.stack([size[:, 0], size[:, 2], size[:, 1]], axis=1)            # local x~l, y~h, z~w
    loc = loc * lhw[:, None, :]
    c, s = np.cos(heading)[:, None], np.sin(heading)[:, None]
    obj = np.stack([c * loc[..., 0] + s * loc[..., 2], loc[..., 1],
                    -s * loc[..., 0] + c * loc[..., 2]], axis=2) + center[:, None, :]
    obj = obj + rng.normal(0, 0.01, size=(B, N, 3))

    # background points inside a pyramidal frustum, depth clamped at 8 m (read3dPoints.m)
    half = np.deg2rad(rng.uniform(10, 25, size=(B, 1)))
    z = rng.uniform(0.5, 8.0, size=(B, N))
    bg = np.stack([z * np.tan(rng.uniform(-1, 1, size=(B, N)) * half),
                   z * np.tan(rng.uniform(-1, 1, size=(B, N)) * half * 0.8), z], axis=2)
    xyz = np.where(seg[..., None], obj, bg)
    feats = rng.uniform(0, 1, size=(B, N, max(C - 3, 0)))
    pc = np.concatenate([xyz, feats], axis=2)[:, :, :C].astype(np.float32)

    # calibration
    tilt = rng.normal(0, np.deg2rad(5.0), size=B)
    ct, st = np.cos(tilt), np.sin(tilt)
    Rtilt = np.zeros((B, 3, 3))
    Rtilt[:, 0, 0] = 1
    Rtilt[:, 1, 1] = ct
    Rtilt[:, 1, 2] = -st
    Rtilt[:, 2, 1] = st
    Rtilt[:, 2, 2] = ct
    dims_choices = np.array([(480, 640), (530, 730), (427, 561)], np.float64)
    img_dim = dims_choices[rng.integers(0, 3, size=B)]                      # (rows, cols)
    f = rng.uniform(520, 580, size=B)
    K = np.zeros((B, 3, 3))
    K[:, 0, 0] = f
    K[:, 1, 1] = f
    K[:, 0, 2] = img_dim[:, 1] / 2 * rng.uniform(0.95, 1.05, size=B)
    K[:, 1, 2] = img_dim[:, 0] / 2 * rng.uniform(0.95, 1.05, size=B)
    K[:, 2, 2] = 1
    rot_frust = rng.uniform(-0.5, 0.5, size=(B, 1))

    # 2D box = bbox of the projected GT box (rotated back out of the frustum frame,
    # tf_util.py:1045-1073) with +-10 % jitter (sunrgbd_data/utils.py:198-211), clipped.
    cr, sr = np.cos(rot_frust[:, 0]), np.sin(rot_frust[:, 0])
    c0 = np.stack([cr * center[:, 0] + sr * center[:, 2], center[:, 1],
                   -sr * center[:, 0] + cr * center[:, 2]], axis=1)
    corners = _box_corners_upright_camera(c0, size, heading + rot_frust[:, 0])
    uv = project_upright_camera_to_image(corners, Rtilt, K)
    lo, hi = uv.min(axis=1), uv.max(axis=1)
    wh = np.maximum(hi - lo, 4.0)
    jit = rng.uniform(-0.1, 0.1, size=(B, 4))
    box2D = np.stack([lo[:, 0] + jit[:, 0] * wh[:, 0], lo[:, 1] + jit[:, 1] * wh[:, 1],
                      hi[:, 0] + jit[:, 2] * wh[:, 0], hi[:, 1] + jit[:, 3] * wh[:, 1]], axis=1)
    box2D[:, 0] = np.clip(box2D[:, 0], 0, img_dim[:, 1] - 2)
    box2D[:, 1] = np.clip(box2D[:, 1], 0, img_dim[:, 0] - 2)
    box2D[:, 2] = np.clip(box2D[:, 2], box2D[:, 0] + 1, img_dim[:, 1])
    box2D[:, 3] = np.clip(box2D[:, 3], box2D[:, 1] + 1, img_dim[:, 0])

    ocls, ores = angle2class(heading)
    one_hot = np.zeros((B, NUM_CLASS), np.float32)
    one_hot[np.arange(B), cls] = 1
    if np.isscalar(is_data_2D):
        is2d = np.full(B, int(is_data_2D), np.int32)
    else:
        is2d = np.asarray(is_data_2D, np.int32)
    out = dict(
        pc=pc, one_hot=one_hot, labels=seg.astype(np.int32),
        centers=center.astype(np.float32),
        y_orient_cls=ocls.astype(np.int32), y_orient_reg=ores.astype(np.float32),
        y_dims_cls=cls.astype(np.int32),
        y_dims_reg=(size - MEAN_DIMS_ARR[cls]).astype(np.float32),
        Rtilt=Rtilt.astype(np.float32), K=K.astype(np.float32),
        rot_frust=rot_frust.astype(np.float32), box2D=box2D.astype(np.float32),
        img_dim=img_dim.astype(np.float32), is_data_2D=is2d)
    # 2D-only samples carry all-zero 3D labels (roi_semi_dataset.py:452-454)
    m = is2d.astype(bool)
    if m.any():
        for k in ('labels', 'centers', 'y_orient_cls', 'y_orient_reg', 'y_dims_cls', 'y_dims_reg'):
            out[k][m] = 0
    return out


def make_boxpc_batch(B, N=2048, C=6, seed=1238, cfg=None):
    """Perturbed (box, point cloud) pairs for the BoxPC-Fit network, SURVEY 8d cfg4:
    half 'fit' band / half 'no-fit' band perturbations in the style of
    box_pc_fit_dataset.py:211-244 (no rejection sampling: y_box_iou is drawn inside the
    band instead of measured, since box_util.box3d_iou is absent from the reference tree)."""
    base = make_batch(B, N, C, seed)
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    fit = rng.random(B) < 0.5
    iou = np.where(fit, rng.uniform(0.7, 1.0, size=B), rng.uniform(0.01, 0.25, size=B))
    mu = np.where(fit, 0.85, 0.13)
    cls = base['y_dims_cls']
    size = MEAN_DIMS_ARR[cls] + base['y_dims_reg']
    dc = rng.uniform(-0.8, 0.8, size=(B, 3)) * (1 - mu)[:, None]
    ds = size * rng.uniform(-0.2, 0.2, size=(B, 3)) * (1 - mu)[:, None]
    da = rng.uniform(0, np.pi, size=B) * (1 - mu)
    heading = class2angle(base['y_orient_cls'], base['y_orient_reg'], to_label_format=False)
    ocls, ores = angle2class(heading + da)
    out = dict(
        pc=base['pc'], one_hot=base['one_hot'], y_seg=base['labels'],
        x_center=(base['centers'] + dc).astype(np.float32),
        x_orient_cls=ocls.astype(np.int32), x_orient_reg=ores.astype(np.float32),
        x_dims_cls=cls.astype(np.int32), x_dims_reg=(base['y_dims_reg'] + ds).astype(np.float32),
        y_box_iou=iou.astype(np.float32), y_center_delta=dc.astype(np.float32),
        y_dims_delta=ds.astype(np.float32), y_orient_delta=da.astype(np.float32))
    return out
