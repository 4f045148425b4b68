"""TEST INFRASTRUCTURE (oracle): numpy float64 restatement of sunrgbd_detection/eval_det.py -- voc_ap (:24-55),
eval_det_cls (:71-157) with get_iou = the oracle's box3d_iou (oracle/box_util.py), eval_det (:159-199).  Pinned
against the reference's own eval_det.py executed here (tests/golden/ref_numpy_helpers.npz, both AP metrics; only
box3d_iou, absent from the reference tree, is shared) and by hand-computed cases in tests/test_oracle_cpu.py.  Only tests/, smoke() and
bench.py's cpu_baseline may import this."""
import numpy as np

from .box_util import box3d_iou


def voc_ap(rec, prec, use_07_metric=False):
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


def eval_det_cls(pred, gt, ovthresh=0.25, use_07_metric=False, return_match=False):
    class_recs, npos = {}, 0
    for img_id in gt.keys():
        bbox = np.array(gt[img_id])
        npos += len(bbox)
        class_recs[img_id] = {'bbox': bbox, 'det': [False] * len(bbox)}
    for img_id in pred.keys():
        if img_id not in gt:
            class_recs[img_id] = {'bbox': np.array([]), 'det': []}
    image_ids, confidence, BB = [], [], []
    for img_id in pred.keys():
        for box, score in pred[img_id]:
            image_ids.append(img_id)
            confidence.append(score)
            BB.append(box)
    confidence = np.array(confidence)
    BB = np.array(BB)
    sorted_ind = np.argsort(-confidence)
    BB = BB[sorted_ind, ...]
    image_ids = [image_ids[x] for x in sorted_ind]
    nd = len(image_ids)
    tp, fp = np.zeros(nd), np.zeros(nd)
    ovs = np.full(nd, -np.inf)
    for d in range(nd):
        R = class_recs[image_ids[d]]
        bb = BB[d, :].astype(float)
        ovmax = -np.inf
        BBGT = R['bbox'].astype(float)
        if BBGT.size > 0:
            for j in range(BBGT.shape[0]):
                iou = box3d_iou(bb, BBGT[j, ...])[0]
                if iou > ovmax:
                    ovmax, jmax = iou, j
        ovs[d] = ovmax
        if ovmax > ovthresh:
            if not R['det'][jmax]:
                tp[d] = 1.
                R['det'][jmax] = 1
            else:
                fp[d] = 1.
        else:
            fp[d] = 1.
    match = (tp.copy(), fp.copy(), ovs)
    fp, tp = np.cumsum(fp), np.cumsum(tp)
    rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    ap = voc_ap(rec, prec, use_07_metric)
    return (rec, prec, ap, match) if return_match else (rec, prec, ap)


def eval_det(pred_all, gt_all, ovthresh={}, use_07_metric=False):
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
