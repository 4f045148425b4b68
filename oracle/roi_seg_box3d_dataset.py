"""TEST INFRASTRUCTURE (oracle): numpy float64 restatement of the label conversions of
sunrgbd_detection/roi_seg_box3d_dataset.py -- rotate_pc_along_y (:37-45), class2angle (:64-71), class2size (:79-82),
from_prediction_to_label_format (:461-466).  Pinned against the reference's own functions executed here
(tests/golden/ref_numpy_helpers.npz, tests/test_oracle_vs_reference_cpu.py) and by hand-computed cases in tests/test_oracle_cpu.py.  Only tests/, smoke() and bench.py's cpu_baseline may import this."""
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


# ------------------------------------------------------------------------------------------------ dataset (labelled layout)
from transferable3d_b200.constants import NUM_CLASS, type2class  # noqa: E402
type2onehotclass = dict(type2class)


def angle2class(angle, num_class):
    """roi_seg_box3d_dataset.py:47-62."""
    angle = angle % (2 * np.pi)
    assert 0 <= angle <= 2 * np.pi
    angle_per_class = 2 * np.pi / float(num_class)
    shifted_angle = (angle + angle_per_class / 2) % (2 * np.pi)
    class_id = int(shifted_angle / angle_per_class)
    return class_id, shifted_angle - (class_id * angle_per_class + angle_per_class / 2)


def size2class(size, type_name):
    return type2class[type_name], size - type_mean_size[type_name]


class ROISegBoxDataset(object):
    """Literal numpy restatement of ROISegBoxDataset.__getitem__ (:259-345, labelled layout) and get_batch (:370-417) over
    already-loaded lists; draws from numpy's global stream exactly like the reference."""

    def __init__(self, lists, npoints, random_flip=False, random_shift=False, rotate_to_center=False, one_hot=False):
        (self.idx_l, self.box2d_l, self.box3d_l, self.image_crop_l, self.points_l, self.label_l, self.cls_type_l, self.heading_l,
         self.size_l, self.rtilt_l, self.k_l, self.frustum_angle_l, self.img_dims_l) = lists
        self.npoints, self.random_flip, self.random_shift = npoints, random_flip, random_shift
        self.rotate_to_center, self.one_hot = rotate_to_center, one_hot

    def get_center_view_rot_angle(self, index):
        return np.pi / 2.0 + self.frustum_angle_l[index]

    def get_box3d_center(self, index):
        return (self.box3d_l[index][0, :] + self.box3d_l[index][6, :]) / 2.0

    def get_center_view_box3d_center(self, index):
        c = (self.box3d_l[index][0, :] + self.box3d_l[index][6, :]) / 2.0
        return rotate_pc_along_y(np.expand_dims(c, 0), self.get_center_view_rot_angle(index)).squeeze()

    def __getitem__(self, index):
        if self.one_hot:
            one_hot_vec = np.zeros((NUM_CLASS))
            one_hot_vec[type2onehotclass[self.cls_type_l[index]]] = 1
        if self.rotate_to_center:
            point_set = rotate_pc_along_y(np.copy(self.points_l[index]), self.get_center_view_rot_angle(index))
        else:
            point_set = self.points_l[index]
        choice = np.random.choice(point_set.shape[0], self.npoints, replace=True)
        point_set = point_set[choice, :]
        rot_angle = self.get_center_view_rot_angle(index)
        box2d, rtilt, k, img_dims = self.box2d_l[index], self.rtilt_l[index], self.k_l[index], self.img_dims_l[index]
        seg = self.label_l[index][choice]
        box3d_center = self.get_center_view_box3d_center(index) if self.rotate_to_center else self.get_box3d_center(index)
        heading_angle = self.heading_l[index] - rot_angle if self.rotate_to_center else self.heading_l[index]
        size_class, size_residual = size2class(self.size_l[index], self.cls_type_l[index])
        if self.random_flip:
            if np.random.random() > 0.5:
                point_set[:, 0] *= -1
                box3d_center[0] *= -1
                heading_angle = np.pi - heading_angle
        if self.random_shift:
            dist = np.sqrt(np.sum(box3d_center[0] ** 2 + box3d_center[1] ** 2))
            shift = np.clip(np.random.randn() * dist * 0.05, dist * 0.8, dist * 1.2)
            point_set[:, 2] += shift
            box3d_center[2] += shift
            height_shift = np.random.random() * 0.4 - 0.2
            point_set[:, 1] += height_shift
            box3d_center[1] += height_shift
        angle_class, angle_residual = angle2class(heading_angle, NUM_HEADING_BIN)
        out = (point_set, None, seg, box3d_center, angle_class, angle_residual, size_class, size_residual, box2d, rtilt, k, rot_angle,
               img_dims)
        return out + ((one_hot_vec,) if self.one_hot else ())

    def get_batch(self, idxs, start_idx, end_idx, num_point, num_channel):
        bsize = end_idx - start_idx
        data, label = np.zeros((bsize, num_point, num_channel)), np.zeros((bsize, num_point), dtype=np.int32)
        center, hcls, hres = np.zeros((bsize, 3)), np.zeros((bsize,), dtype=np.int32), np.zeros((bsize,))
        scls, sres = np.zeros((bsize,), dtype=np.int32), np.zeros((bsize, 3))
        box2d, rtilts, ks = np.zeros((bsize, 4)), np.zeros((bsize, 3, 3)), np.zeros((bsize, 3, 3))
        rot, img_dims, one_hot = np.zeros((bsize,)), np.zeros((bsize, 2)), np.zeros((bsize, NUM_CLASS))
        for i in range(bsize):
            item = self[idxs[i + start_idx]]
            ps, _, seg, c, hc, hr, sc, sr, b2, rt_, k, ra, idm = item[:13]
            if self.one_hot:
                one_hot[i] = item[13]
            data[i, ...] = ps[:, 0:num_channel]
            label[i, :], center[i, :], hcls[i], hres[i], scls[i], sres[i] = seg, c, hc, hr, sc, sr
            box2d[i], rtilts[i], ks[i], rot[i], img_dims[i] = b2, rt_, k, ra, idm
        out = (data, None, label, center, hcls, hres, scls, sres, box2d, rtilts, ks, rot, img_dims)
        return out + ((one_hot,) if self.one_hot else ())
