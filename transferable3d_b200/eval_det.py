"""Mirror of sunrgbd/sunrgbd_detection/eval_det.py (voc_ap, get_iou, eval_det_cls, eval_det: same names, arguments and
returned (rec, prec, ap)) and of evaluate.evaluate_predictions' box construction (evaluate.py:53-67).

The per-detection python loop of eval_det_cls (:118-145: 3D IoU against every ground-truth box of the image, greedy
first-claim matching) runs as one kernel per class (t3d_det_match, one thread per image); sorting by score, the two
cumulative sums and voc_ap stay host numpy exactly as in the reference (O(nd) bookkeeping).
"""
import ctypes

import numpy as np
import torch

from . import runtime as rt
from ._lib import ptr, stream, call, t3d_det_match_args
from . import box_util


def voc_ap(rec, prec, use_07_metric=False):
    """eval_det.py:24-55."""
    if use_07_metric:
        ap = 0.
        for t in np.arange(0., 1.1, 0.1):
            p = 0 if np.sum(rec >= t) == 0 else np.max(prec[rec >= t])
            ap = ap + p / 11.
        return ap
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], prec, [0.]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def get_iou(bb1, bb2):
    """eval_det.py:63-69: 3D IoU of two (8,3) corner sets."""
    i3, _ = box_util.box3d_iou(np.asarray(bb1, dtype=np.float32).reshape(1, 8, 3), np.asarray(bb2, dtype=np.float32).reshape(1, 8, 3))
    return float(i3[0])


def match_detections(pred, gt, ovthresh=0.25, device=None):
    """The sort + matching part of eval_det_cls -> dict(tp, fp, ovmax, jmax (numpy, in sorted order), npos, sorted_ind)."""
    dev = torch.device(device) if device is not None else rt.default_device()
    img_ids = list(gt.keys()) + [i for i in pred.keys() if i not in gt]          # class_recs order (eval_det.py:86-96)
    index = {img_id: n for n, img_id in enumerate(img_ids)}
    npos = sum(len(gt[i]) for i in gt)
    image_ids, confidence, BB = [], [], []
    for img_id in pred.keys():
        for box, score in pred[img_id]:
            image_ids.append(index[img_id])
            confidence.append(score)
            BB.append(box)
    nd = len(image_ids)
    if nd == 0:
        return dict(tp=np.zeros(0), fp=np.zeros(0), ovmax=np.zeros(0), jmax=np.zeros(0, dtype=np.int64), npos=npos,
                    sorted_ind=np.zeros(0, dtype=np.int64))
    confidence = np.array(confidence)
    sorted_ind = np.argsort(-confidence)
    BB = np.asarray(BB, dtype=np.float32).reshape(nd, 8, 3)[sorted_ind]
    det_img = np.asarray(image_ids, dtype=np.int64)[sorted_ind]
    nimg = len(img_ids)
    order = np.argsort(det_img, kind='stable')                                    # positions grouped by image, score order kept
    det_off = np.zeros(nimg + 1, dtype=np.int32)
    np.cumsum(np.bincount(det_img, minlength=nimg), out=det_off[1:])
    gt_list = [np.asarray(gt[i], dtype=np.float32).reshape(-1, 8, 3) if i in gt and len(gt[i]) else np.zeros((0, 8, 3), np.float32)
               for i in img_ids]
    gt_off = np.zeros(nimg + 1, dtype=np.int32)
    np.cumsum([g.shape[0] for g in gt_list], out=gt_off[1:])
    gt_all = np.concatenate(gt_list, 0) if gt_list else np.zeros((0, 8, 3), np.float32)
    ng = int(gt_all.shape[0])
    T = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(device=dev, dtype=dt)
    d_bb, d_off, d_idx = T(BB, torch.float32), T(det_off, torch.int32), T(order.astype(np.int32), torch.int32)
    d_gt = T(gt_all if ng else np.zeros((1, 8, 3), np.float32), torch.float32)
    d_goff = T(gt_off, torch.int32)
    tp = torch.empty(nd, dtype=torch.float32, device=dev)
    fp = torch.empty_like(tp)
    ov = torch.empty_like(tp)
    jm = torch.empty(nd, dtype=torch.int32, device=dev)
    scratch = torch.empty(max(ng, 1), dtype=torch.uint8, device=dev)
    a = t3d_det_match_args(ptr(d_bb), ptr(d_off), ptr(d_idx), ptr(d_gt), ptr(d_goff), nimg, nd, ng, float(ovthresh),
                           ptr(tp), ptr(fp), ptr(ov), ptr(jm), ptr(scratch))
    call('t3d_det_match', ctypes.byref(a), stream())
    return dict(tp=tp.cpu().numpy().astype(np.float64), fp=fp.cpu().numpy().astype(np.float64), ovmax=ov.cpu().numpy(),
                jmax=jm.cpu().numpy(), npos=npos, sorted_ind=sorted_ind)


def eval_det_cls(pred, gt, ovthresh=0.25, use_07_metric=False):
    """eval_det.py:71-157: pred {img_id: [(bbox (8,3), score)]}, gt {img_id: [bbox]} -> (rec, prec, ap)."""
    m = match_detections(pred, gt, ovthresh)
    fp = np.cumsum(m['fp'])
    tp = np.cumsum(m['tp'])
    rec = tp / float(m['npos'])
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return rec, prec, voc_ap(rec, prec, use_07_metric)


def eval_det(pred_all, gt_all, ovthresh={}, use_07_metric=False):
    """eval_det.py:159-199: pred_all {img_id: [(classname, bbox, score)]}, gt_all {img_id: [(classname, bbox)]}."""
    pred, gt = {}, {}
    for img_id in pred_all.keys():
        for classname, bbox, score in pred_all[img_id]:
            pred.setdefault(classname, {}).setdefault(img_id, [])
            gt.setdefault(classname, {}).setdefault(img_id, [])
            pred[classname][img_id].append((bbox, score))
    for img_id in gt_all.keys():
        for classname, bbox in gt_all[img_id]:
            gt.setdefault(classname, {}).setdefault(img_id, []).append(bbox)
    rec, prec, ap = {}, {}, {}
    for classname in gt.keys():
        thresh = ovthresh[classname] if type(ovthresh) is dict else ovthresh
        rec[classname], prec[classname], ap[classname] = eval_det_cls(pred.get(classname, {}), gt[classname], thresh, use_07_metric)
    return rec, prec, ap


def prediction_corners(center_list, heading_cls_list, heading_res_list, size_cls_list, size_res_list, rot_angle_list):
    """evaluate.evaluate_predictions' predicted boxes (evaluate.py:53-67), batched on the device: class2angle / class2size
    -> get_3d_box -> rotate_pc_along_y(corners, -rot_angle).  Returns (B,8,3) numpy float32."""
    from .constants import NUM_HEADING_BIN, MEAN_DIMS_ARR
    center = np.asarray(center_list, dtype=np.float64).reshape(-1, 3)
    hc, sc = np.asarray(heading_cls_list).reshape(-1), np.asarray(size_cls_list).reshape(-1)
    ang = hc * (2 * np.pi / NUM_HEADING_BIN) + np.asarray(heading_res_list, dtype=np.float64).reshape(-1)
    ang = np.where(ang > np.pi, ang - 2 * np.pi, ang)
    size = MEAN_DIMS_ARR[sc] + np.asarray(size_res_list, dtype=np.float64).reshape(-1, 3)
    corners = box_util.get_3d_box(size.astype(np.float32), ang.astype(np.float32), center.astype(np.float32))
    rot = torch.as_tensor(-np.asarray(rot_angle_list, dtype=np.float32).reshape(-1)).to(corners.device)
    c, s = torch.cos(rot)[:, None], torch.sin(rot)[:, None]
    x, z = corners[:, :, 0].clone(), corners[:, :, 2].clone()
    corners[:, :, 0] = c * x - s * z                      # rotate_pc_along_y: [x, z] . [[c, -s], [s, c]]^T
    corners[:, :, 2] = s * x + c * z
    return corners.cpu().numpy()
