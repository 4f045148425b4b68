"""TEST INFRASTRUCTURE (oracle): numpy float64 restatement of the label conversions of
sunrgbd_detection/roi_seg_box3d_dataset.py -- rotate_pc_along_y (:37-45), class2angle (:64-71), class2size (:79-82),
from_prediction_to_label_format (:461-466).  Parity unpinned (the reference ships no vectors); pinned here by
hand-computed cases in tests/test_oracle_cpu.py.  Only tests/, smoke() and bench.py's cpu_baseline may import this."""
import numpy as np

from transferable3d_b200.constants import NUM_HEADING_BIN, type_mean_size, class2type


def rotate_pc_along_y(pc, rot_angle):
    cosval, sinval = np.cos(rot_angle), np.sin(rot_angle)
    rotmat = np.array([[cosval, -sinval], [sinval, cosval]])
    pc = np.array(pc, dtype=np.float64)
    pc[:, [0, 2]] = np.dot(pc[:, [0, 2]], np.transpose(rotmat))
    return pc


def class2angle(pred_cls, residual, num_class, to_label_format=True):
    angle = pred_cls * (2 * np.pi / float(num_class)) + residual
    if to_label_format and angle > np.pi:
        angle = angle - 2 * np.pi
    return angle


def class2size(pred_cls, residual):
    return type_mean_size[class2type[int(pred_cls)]] + residual


def from_prediction_to_label_format(center, angle_class, angle_res, size_class, size_res, rot_angle):
    l, w, h = class2size(size_class, np.asarray(size_res, dtype=np.float64))
    ry = class2angle(angle_class, angle_res, NUM_HEADING_BIN) + rot_angle
    tx, ty, tz = rotate_pc_along_y(np.expand_dims(np.asarray(center, dtype=np.float64), 0), -rot_angle).squeeze()
    ty += h / 2.0
    return h, w, l, tx, ty, tz, ry
