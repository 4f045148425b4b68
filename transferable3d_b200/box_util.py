"""GPU stand-in for the `box_util` module the reference imports but does not ship (roi_seg_box3d_dataset.py:15,
box_pc_fit_dataset.py:17; it is train/box_util.py of charlesq34/frustum-pointnets) and for its callers on the BoxPC-Fit
data path: get_3d_box / compute_box3d_iou (roi_seg_box3d_dataset.py:84-139), get_box3d_iou and
BoxPCFitDataset.perturb_box_to_diff_ious (box_pc_fit_dataset.py:35-42, 211-244).  Batched over boxes: every function
takes (B, ...) tensors where the reference loops over one box at a time in python; kernels in csrc/box_ops.cuh."""
import ctypes

import numpy as np
import torch

from . import runtime as rt
from ._lib import ptr, stream, call, t3d_compute_iou_args, t3d_perturb_args
from .constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER, MEAN_DIMS_ARR


_MS = {}


def _mean_size(dev):
    key = str(dev)
    if key not in _MS:
        _MS[key] = torch.as_tensor(np.asarray(MEAN_DIMS_ARR, dtype=np.float32)).to(dev).contiguous()
    return _MS[key]


def _f(t, dev=None):
    t = t if torch.is_tensor(t) else torch.as_tensor(np.asarray(t))
    return t.to(device=dev or (t.device if t.is_cuda else 'cuda'), dtype=torch.float32).contiguous()


def get_3d_box(box_size, heading_angle, center):
    """roi_seg_box3d_dataset.py:84-100, batched: (B,3) sizes (l,w,h), (B,) headings, (B,3) centres -> (B,8,3) corners."""
    box_size, heading_angle, center = _f(box_size), _f(heading_angle), _f(center)
    B = box_size.shape[0]
    out = torch.empty((B, 8, 3), dtype=torch.float32, device=box_size.device)
    call('t3d_get_3d_box', ptr(box_size), ptr(heading_angle), ptr(center), B, ptr(out), stream())
    return out


def box3d_iou(corners1, corners2):
    """box_util.box3d_iou, batched: (B,8,3) x (B,8,3) -> (iou_3d (B,), iou_2d (B,))."""
    corners1, corners2 = _f(corners1), _f(corners2)
    B = corners1.shape[0]
    i3 = torch.empty((B,), dtype=torch.float32, device=corners1.device)
    i2 = torch.empty_like(i3)
    call('t3d_box3d_iou', ptr(corners1), ptr(corners2), B, ptr(i3), ptr(i2), stream())
    return i3, i2


def get_box3d_iou(center_A, box_size_A, heading_angle_A, center_B, box_size_B, heading_angle_B):
    """box_pc_fit_dataset.py:35-39, batched."""
    return box3d_iou(get_3d_box(box_size_A, heading_angle_A, center_A), get_3d_box(box_size_B, heading_angle_B, center_B))


def compute_box3d_iou(center_pred, heading_logits, heading_residuals, size_logits, size_residuals, center_label,
                      heading_class_label, heading_residual_label, size_class_label, size_residual_label):
    """roi_seg_box3d_dataset.py:102-139 in one kernel -> (iou2ds (B,), iou3ds (B,))."""
    dev = center_pred.device if torch.is_tensor(center_pred) and center_pred.is_cuda else torch.device('cuda')
    F = lambda t: _f(t, dev)
    I = lambda t: (t if torch.is_tensor(t) else torch.as_tensor(np.asarray(t))).to(device=dev, dtype=torch.int32).contiguous()
    cp, hl, hr, sl, sr = F(center_pred), F(heading_logits), F(heading_residuals), F(size_logits), F(size_residuals)
    cl, hcl, hrl, scl, srl = F(center_label), I(heading_class_label), F(heading_residual_label), I(size_class_label), F(size_residual_label)
    B = cp.shape[0]
    ms = _mean_size(dev)
    i2 = torch.empty((B,), dtype=torch.float32, device=dev)
    i3 = torch.empty_like(i2)
    a = t3d_compute_iou_args(ptr(cp), ptr(hl), ptr(hr), ptr(sl), ptr(sr), ptr(cl), ptr(hcl), ptr(hrl), ptr(scl), ptr(srl), ptr(ms),
                             B, hl.shape[1], sl.shape[1], ptr(i2), ptr(i3))
    call('t3d_compute_box3d_iou', ctypes.byref(a), stream())
    return i2, i3


def perturb_box_to_diff_ious(box3d_center, size, heading_angle, iou_bounds, center_perturbation=0.8, size_perturbation=0.2,
                             angle_perturbation=np.pi, seed=0, max_attempts=100000):
    """BoxPCFitDataset.perturb_box_to_diff_ious (box_pc_fit_dataset.py:211-244), batched: one rejection-sampling loop per
    box on the GPU ('philox' stream of oracle/box_pc_fit_dataset.py).  iou_bounds: (B,2) or one (lo, hi) pair.
    -> (new_center, new_size, new_heading, iou3d, y_center_delta, y_size_delta, y_angle_delta, attempts)."""
    c, s, h = _f(box3d_center), _f(size), _f(heading_angle)
    dev, B = c.device, c.shape[0]
    b = _f(iou_bounds, dev)
    if b.dim() == 1:
        b = b.reshape(1, 2).expand(B, 2).contiguous()
    E = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    nc, ns, nh, iou, dc, ds, da = E(B, 3), E(B, 3), E(B), E(B), E(B, 3), E(B, 3), E(B)
    att = torch.empty((B,), dtype=torch.int32, device=dev)
    a = t3d_perturb_args(ptr(c), ptr(s), ptr(h), ptr(b), B, int(max_attempts), float(center_perturbation), float(size_perturbation),
                         float(angle_perturbation), int(seed), ptr(nc), ptr(ns), ptr(nh), ptr(iou), ptr(dc), ptr(ds), ptr(da), ptr(att))
    call('t3d_perturb_boxes', ctypes.byref(a), stream())
    return nc, ns, nh, iou, dc, ds, da, att
