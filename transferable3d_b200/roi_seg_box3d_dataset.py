"""Label <-> prediction conversions of sunrgbd_detection/roi_seg_box3d_dataset.py (same names and argument order).

angle2class / size2class / class2angle / class2size are the dataset's scalar label encoders (:47-82, host logic in the
reference too).  from_prediction_to_label_format (:461-466) runs on the device for a whole batch
(t3d_prediction_to_label); the scalar reference signature is the B = 1 case of the same kernel.
"""
import numpy as np
import torch

from . import runtime as rt
from ._lib import call, ptr, stream
from .constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER, MEAN_DIMS_ARR, type2class, class2type, type_mean_size  # noqa: F401


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
