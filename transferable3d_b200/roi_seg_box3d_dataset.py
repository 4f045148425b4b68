"""Label <-> prediction conversions of sunrgbd_detection/roi_seg_box3d_dataset.py (same names and argument order).

angle2class / size2class / class2angle / class2size are the dataset's scalar label encoders (:47-82, host logic in the
reference too).  from_prediction_to_label_format (:461-466) runs on the device for a whole batch
(t3d_prediction_to_label); the scalar reference signature is the B = 1 case of the same kernel.
"""
import numpy as np
import torch

from . import runtime as rt
from ._lib import call, ptr, stream
from .constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER, NUM_CLASS, MEAN_DIMS_ARR, type2class, class2type, type_mean_size  # noqa: F401


def angle2class(angle, num_class):
    """roi_seg_box3d_dataset.py:47-62."""
    angle = angle % (2 * np.pi)
    assert 0 <= angle <= 2 * np.pi
    angle_per_class = 2 * np.pi / float(num_class)
    shifted_angle = (angle + angle_per_class / 2) % (2 * np.pi)
    class_id = int(shifted_angle / angle_per_class)
    residual_angle = shifted_angle - (class_id * angle_per_class + angle_per_class / 2)
    return class_id, residual_angle


def class2angle(pred_cls, residual, num_class, to_label_format=True):
    """roi_seg_box3d_dataset.py:64-71."""
    angle = pred_cls * (2 * np.pi / float(num_class)) + residual
    if to_label_format and angle > np.pi:
        angle = angle - 2 * np.pi
    return angle


def size2class(size, type_name):
    """roi_seg_box3d_dataset.py:73-77."""
    return type2class[type_name], size - type_mean_size[type_name]


def class2size(pred_cls, residual):
    """roi_seg_box3d_dataset.py:79-82."""
    return type_mean_size[class2type[pred_cls]] + residual


def from_prediction_to_label_format_batch(center, angle_class, angle_res, size_class, size_res, rot_angle, device=None):
    """(B,3), (B,), (B,), (B,), (B,3), (B,) -> (B,7) device tensor of (h, w, l, tx, ty, tz, ry) per box."""
    dev = torch.device(device) if device is not None else (center.device if torch.is_tensor(center) and center.is_cuda
                                                           else rt.default_device())
    F = lambda v: torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).to(device=dev, dtype=torch.float32).contiguous()
    I = lambda v: torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).to(device=dev, dtype=torch.int32).contiguous()
    center, angle_res, size_res, rot_angle = F(center).view(-1, 3), F(angle_res).view(-1), F(size_res).view(-1, 3), F(rot_angle).view(-1)
    angle_class, size_class = I(angle_class).view(-1), I(size_class).view(-1)
    B = center.shape[0]
    mean_size = F(MEAN_DIMS_ARR)
    out = torch.empty((B, 7), dtype=torch.float32, device=dev)
    call('t3d_prediction_to_label', ptr(center), ptr(angle_class), ptr(angle_res), ptr(size_class), ptr(size_res), ptr(rot_angle),
         ptr(mean_size), B, NUM_HEADING_BIN, ptr(out), stream())
    return out


def from_prediction_to_label_format(center, angle_class, angle_res, size_class, size_res, rot_angle):
    """roi_seg_box3d_dataset.py:461-466 -> (h, w, l, tx, ty, tz, ry) python floats."""
    out = from_prediction_to_label_format_batch(np.asarray(center).reshape(1, 3), [angle_class], [angle_res], [size_class],
                                                np.asarray(size_res).reshape(1, 3), [rot_angle])
    return tuple(float(v) for v in out[0].cpu())


# ------------------------------------------------------------------------------------------------ dataset + batch assembly
def rotate_pc_along_y(pc, rot_angle):
    """roi_seg_box3d_dataset.py:37-45 (host numpy, in place like the reference)."""
    cosval, sinval = np.cos(rot_angle), np.sin(rot_angle)
    rotmat = np.array([[cosval, -sinval], [sinval, cosval]])
    pc[:, [0, 2]] = np.dot(pc[:, [0, 2]], np.transpose(rotmat))
    return pc


class ROISegBoxDataset(object):
    """roi_seg_box3d_dataset.ROISegBoxDataset (:195-459) with the same constructor, list attributes (idx_l, box2d_l, ...,
    frustum_angle_l) and get_batch signature / tuple order.  The point sets also live on the device as one flat array, and
    get_batch assembles a whole batch in one kernel (t3d_assemble_frustum_batch) instead of a python loop over
    __getitem__; it returns device tensors (the image slot, all zeros and unused in the reference, is None).
    Random draws follow the reference's per-item order on numpy's global stream (choice, then flip, then the two shift
    draws), so a seeded run selects the same points as the reference would."""

    def __init__(self, classes, npoints, split, classes_to_drop=[], classes_to_drop_prob=0, random_flip=False, random_shift=False,
                 rotate_to_center=False, overwritten_data_path=None, from_rgb_detection=False, one_hot=False, device=None):
        from .utils import load_zipped_pickle
        self.classes, self.npoints = classes, npoints
        self.random_flip, self.random_shift, self.rotate_to_center = random_flip, random_shift, rotate_to_center
        self.one_hot, self.from_rgb_detection = one_hot, from_rgb_detection
        assert overwritten_data_path is not None, 'pass the prepared frustum file (sunrgbd_data.py output)'
        lists = load_zipped_pickle(overwritten_data_path)
        if from_rgb_detection:
            names = ('idx_l', 'box2d_l', 'image_crop_l', 'points_l', 'cls_type_l', 'frustum_angle_l', 'prob_l')
        else:
            names = ('idx_l', 'box2d_l', 'box3d_l', 'image_crop_l', 'points_l', 'label_l', 'cls_type_l', 'heading_l', 'size_l',
                     'rtilt_l', 'k_l', 'frustum_angle_l', 'img_dims_l')
        assert len(lists) == len(names), 'expected the %d-list layout, got %d lists' % (len(names), len(lists))
        for n in names:
            setattr(self, n, [])
        ci = names.index('cls_type_l')
        for rec in zip(*lists):
            if rec[ci] not in self.classes:
                continue
            if (not from_rgb_detection) and rec[ci] in classes_to_drop and (np.random.rand() < classes_to_drop_prob):
                continue
            for n, v in zip(names, rec):
                getattr(self, n).append(v)
        self.device = torch.device(device) if device is not None else rt.default_device()
        self._upload()

    def _upload(self):
        dev, F = self.device, len(self.points_l)
        T = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
        counts = np.array([p.shape[0] for p in self.points_l], dtype=np.int64)
        self._off_host = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self._C = int(self.points_l[0].shape[1]) if F else 6
        self._points = T(np.concatenate(self.points_l, 0) if F else np.zeros((1, self._C)), torch.float32)
        self._off = T(self._off_host, torch.int64)
        self._fangle = T(np.asarray(self.frustum_angle_l, dtype=np.float64).reshape(-1), torch.float32)
        self._cls = T([type2class[t] for t in self.cls_type_l], torch.int32)
        self._mean = T(MEAN_DIMS_ARR, torch.float32)
        if not self.from_rgb_detection:
            self._labels = T(np.concatenate(self.label_l, 0), torch.int32)
            self._box3d = T(np.asarray(self.box3d_l, dtype=np.float64).reshape(F, 8, 3), torch.float32)
            self._heading = T(np.asarray(self.heading_l, dtype=np.float64).reshape(-1), torch.float32)
            self._size = T(np.asarray(self.size_l, dtype=np.float64).reshape(F, 3), torch.float32)
            self._box2d = T(np.asarray(self.box2d_l, dtype=np.float64).reshape(F, 4), torch.float32)
            self._rtilt = T(np.asarray(self.rtilt_l, dtype=np.float64).reshape(F, 3, 3), torch.float32)
            self._k = T(np.asarray(self.k_l, dtype=np.float64).reshape(F, 3, 3), torch.float32)
            self._img_dims = T(np.asarray(self.img_dims_l, dtype=np.float64).reshape(F, 2), torch.float32)
        else:
            self._prob = T(np.asarray(self.prob_l, dtype=np.float64).reshape(-1), torch.float32)

    def __len__(self):
        return len(self.points_l)

    def get_center_view_rot_angle(self, index):
        return np.pi / 2.0 + self.frustum_angle_l[index]

    def get_box3d_center(self, index):
        return (self.box3d_l[index][0, :] + self.box3d_l[index][6, :]) / 2.0

    def get_center_view_box3d_center(self, index):
        c = (self.box3d_l[index][0, :] + self.box3d_l[index][6, :]) / 2.0
        return rotate_pc_along_y(np.expand_dims(c, 0), self.get_center_view_rot_angle(index)).squeeze()

    def get_center_view_box3d(self, index):
        return rotate_pc_along_y(np.copy(self.box3d_l[index]), self.get_center_view_rot_angle(index))

    def get_center_view_point_set(self, index):
        return rotate_pc_along_y(np.copy(self.points_l[index]), self.get_center_view_rot_angle(index))

    def _draws(self, sel):
        """numpy's global stream in the order of __getitem__ (:272, :319-330), item by item."""
        B, N = len(sel), self.npoints
        choice = np.empty((B, N), dtype=np.int32)
        flip = np.zeros(B, dtype=np.uint8)
        sz, sy = np.zeros(B, dtype=np.float32), np.zeros(B, dtype=np.float32)
        labelled = not self.from_rgb_detection
        for i, f in enumerate(sel):
            choice[i] = np.random.choice(self.points_l[f].shape[0], N, replace=True)
            if not labelled:
                continue
            if self.random_flip or self.random_shift:
                c = self.get_center_view_box3d_center(f) if self.rotate_to_center else self.get_box3d_center(f).copy()
            if self.random_flip and np.random.random() > 0.5:
                flip[i] = 1
                c[0] *= -1
            if self.random_shift:
                dist = np.sqrt(np.sum(c[0] ** 2 + c[1] ** 2))
                sz[i] = np.clip(np.random.randn() * dist * 0.05, dist * 0.8, dist * 1.2)
                sy[i] = np.random.random() * 0.4 - 0.2
        return choice, flip, sz, sy

    def get_batch(self, idxs, start_idx, end_idx, num_point, num_channel, from_rgb_detection=False):
        """:370-459 -> the reference's tuple, as device tensors:
        labelled: (data (B,N,C), None, label (B,N) i32, center (B,3), heading_class, heading_residual, size_class,
                   size_residual (B,3), box2d (B,4), rtilts (B,3,3), ks (B,3,3), rot_angle (B,), img_dims (B,2) [, one_hot (B,10)])
        rgb detection: (data, None, rot_angle, prob [, one_hot], oracle_y_seg zeros)."""
        import ctypes
        from ._lib import t3d_assemble_args, load, check
        assert num_point == self.npoints and from_rgb_detection == self.from_rgb_detection
        dev = self.device
        sel = np.asarray([idxs[i + start_idx] for i in range(end_idx - start_idx)], dtype=np.int32)
        B, N = len(sel), num_point
        choice, flip, sz, sy = self._draws(sel)
        T = lambda a, dt: torch.as_tensor(a).to(device=dev, dtype=dt)
        d_sel, d_choice = T(sel, torch.int32), T(choice, torch.int32)
        aug = (not self.from_rgb_detection) and (self.random_flip or self.random_shift)
        d_flip, d_sz, d_sy = (T(flip, torch.uint8), T(sz, torch.float32), T(sy, torch.float32)) if aug else (None, None, None)
        E = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt, device=dev)
        data, rot = E(B, N, num_channel), E(B)
        labelled = not self.from_rgb_detection
        label = E(B, N, dt=torch.int32) if labelled else None
        center, hcls, hres, scls, sres = (E(B, 3), E(B, dt=torch.int32), E(B), E(B, dt=torch.int32), E(B, 3)) if labelled else (None,) * 5
        a = t3d_assemble_args(ptr(self._points), self._C, ptr(self._labels) if labelled else None, ptr(self._off), ptr(d_sel), ptr(d_choice),
                              ptr(self._fangle), ptr(self._box3d) if labelled else None, ptr(self._heading) if labelled else None,
                              ptr(self._size) if labelled else None, ptr(self._cls), ptr(self._mean), ptr(d_flip), ptr(d_sz), ptr(d_sy),
                              B, N, int(num_channel), int(bool(self.rotate_to_center)), NUM_HEADING_BIN, ptr(data), ptr(label),
                              ptr(center), ptr(hcls), ptr(hres), ptr(scls), ptr(sres), ptr(rot))
        check(load().t3d_assemble_frustum_batch(ctypes.byref(a), stream()))
        lsel = d_sel.long()
        one_hot = torch.nn.functional.one_hot(self._cls[lsel].long(), NUM_CLASS).to(torch.float32) if self.one_hot else None
        if not labelled:
            out = (data, None, rot, self._prob[lsel])
            return out + ((one_hot,) if self.one_hot else ()) + (torch.zeros((B, N), dtype=torch.float32, device=dev),)
        out = (data, None, label, center, hcls, hres, scls, sres, self._box2d[lsel], self._rtilt[lsel], self._k[lsel], rot,
               self._img_dims[lsel])
        return out + ((one_hot,) if self.one_hot else ())


def get_3d_box(box_size, heading_angle, center):
    """roi_seg_box3d_dataset.py:84-100 (the training scripts import it from this module: train_boxpc.py:29); kernel in box_util."""
    from . import box_util
    return box_util.get_3d_box(box_size, heading_angle, center)


def compute_box3d_iou(center_pred, heading_logits, heading_residuals, size_logits, size_residuals, center_label,
                      heading_class_label, heading_residual_label, size_class_label, size_residual_label):
    """roi_seg_box3d_dataset.py:102-139 (imported from this module by semisup_v1_sunrgbd.py:23); kernel in box_util."""
    from . import box_util
    return box_util.compute_box3d_iou(center_pred, heading_logits, heading_residuals, size_logits, size_residuals, center_label,
                                      heading_class_label, heading_residual_label, size_class_label, size_residual_label)
