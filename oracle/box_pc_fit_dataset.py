"""TEST INFRASTRUCTURE (oracle): BoxPCFitDataset.perturb_box_to_diff_ious (sunrgbd_detection/box_pc_fit_dataset.py:211-244),
the rejection sampler that perturbs a ground-truth box until its 3D IoU with the original lies inside a band.

The reference draws from numpy's global MT19937 stream, box after box (inherently serial).  Two modes, as for the
512-point resampling (oracle/model_util.py): `numpy_legacy` = the literal draws; `philox` = the same algorithm on a
counter-based stream so that a GPU thread per box reproduces it: attempt t of box b uses
Philox4x32-10(counter = (t, block, b, 0), key = seed); block 0: words 0..2 -> centre deltas, word 3 -> angle delta;
block 1: words 0..2 -> size deltas; u = (word >> 8) * 2^-24."""
import numpy as np

from .model_util import philox4x32_10
from .box_util import get_box3d_iou


def inrange(val, low, high):
    """box_pc_fit_dataset.py:41-42 (strict)."""
    return (val > low) and (val < high)


def _u01(w):
    return (np.uint32(w) >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0)


def perturb_box_to_diff_ious(box3d_center, size, heading_angle, iou_bounds, center_perturbation=0.8, size_perturbation=0.2,
                             angle_perturbation=np.pi, rng_mode='numpy_legacy', rng=None, seed=0, box_index=0, max_attempts=100000):
    """-> (new_center, new_size, new_heading, iou3d, y_center_delta, y_size_delta, y_angle_delta, attempts)."""
    box3d_center, size = np.asarray(box3d_center, dtype=np.float64), np.asarray(size, dtype=np.float64)
    iou_mean = np.mean(iou_bounds)
    cp = center_perturbation * (1 - iou_mean)
    sp = size_perturbation * (1 - iou_mean)
    ap = angle_perturbation * (1 - iou_mean)
    rng = rng if rng is not None else np.random
    iou3d, count = -1, 0
    while not inrange(iou3d, iou_bounds[0], iou_bounds[1]):
        if count >= max_attempts:
            count = -1
            break
        if rng_mode == 'numpy_legacy':
            y_center_delta = rng.uniform(-cp, cp, size=3)
            y_size_delta = np.multiply(size, rng.uniform(-sp, +sp, size=3))
            y_angle_delta = rng.uniform(0, ap)
        else:
            k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
            r0 = [int(w) for w in philox4x32_10(np.uint32(count), np.uint32(0), np.uint32(box_index), np.uint32(0), k0, k1)]
            r1 = [int(w) for w in philox4x32_10(np.uint32(count), np.uint32(1), np.uint32(box_index), np.uint32(0), k0, k1)]
            y_center_delta = np.array([-cp + 2 * cp * _u01(r0[k]) for k in range(3)])
            y_angle_delta = ap * _u01(r0[3])
            y_size_delta = np.array([size[k] * (-sp + 2 * sp * _u01(r1[k])) for k in range(3)])
        new_box3d_center = box3d_center + y_center_delta
        new_size = size + y_size_delta
        new_heading_angle = heading_angle + y_angle_delta
        iou3d, _ = get_box3d_iou(box3d_center, size, heading_angle, new_box3d_center, new_size, new_heading_angle)
        assert 0. <= iou3d <= 1. + 1e-9
        count += 1
    return new_box3d_center, new_size, new_heading_angle, iou3d, y_center_delta, y_size_delta, y_angle_delta, count
