"""-m gpu: detection evaluation -- eval_det / eval_det_cls with the matching loop on the device
(t3d_det_match) against the numpy float64 restatement of sunrgbd_detection/eval_det.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(seed, nimg=60, classes=('bed', 'chair', 'table')):
    """Synthetic detections: per image a few GT boxes per class; predictions = jittered copies (hits of varying quality,
    duplicates, misses) plus images with predictions only and GT only."""
    from oracle import box_util as ob
    from transferable3d_b200.constants import type_mean_size
    rng = np.random.RandomState(seed)
    pred_all, gt_all = {}, {}
    for img in range(nimg):
        gts, preds = [], []
        for _ in range(rng.randint(0, 5)):
            cls = classes[rng.randint(len(classes))]
            size = type_mean_size[cls] * rng.uniform(0.8, 1.2, 3)
            center = np.array([rng.uniform(-3, 3), rng.uniform(-0.5, 0.5), rng.uniform(1.5, 6)])
            heading = rng.uniform(-np.pi, np.pi)
            gts.append((cls, ob.get_3d_box(size, heading, center)))
            for _ in range(rng.randint(0, 4)):         # 0-3 detections around this box
                j = rng.choice([0.02, 0.1, 0.3, 0.6])
                c2 = center + rng.randn(3) * j * size
                s2 = size * (1 + rng.randn(3) * 0.5 * j)
                preds.append((cls, ob.get_3d_box(np.abs(s2) + 0.05, heading + rng.randn() * j, c2), float(rng.rand())))
        for _ in range(rng.randint(0, 2)):             # stray detections
            cls = classes[rng.randint(len(classes))]
            preds.append((cls, ob.get_3d_box(type_mean_size[cls], rng.uniform(-3, 3), rng.uniform(-5, 5, 3) + [0, 0, 5]), float(rng.rand())))
        if img % 11 != 3:
            gt_all[img] = gts
        if preds and img % 7 != 5:
            pred_all[img] = preds
    return pred_all, gt_all


@pytest.mark.parametrize('seed', [0, 1])
def test_eval_det_vs_oracle(seed):
    from transferable3d_b200 import eval_det as ed
    from oracle import eval_det as o
    pred_all, gt_all = _scene(seed)
    rec, prec, ap = ed.eval_det(pred_all, gt_all, 0.25)
    orec, oprec, oap = o.eval_det(pred_all, gt_all, 0.25)
    assert set(ap) == set(oap) and len(ap) == 3
    for c in oap:
        assert rec[c].shape == orec[c].shape
        assert np.array_equal(rec[c], orec[c]) and np.array_equal(prec[c], oprec[c]), c      # identical tp / fp sequences
        assert abs(ap[c] - oap[c]) < 1e-12
    rec7, prec7, ap7 = ed.eval_det(pred_all, gt_all, {'bed': 0.25, 'chair': 0.5, 'table': 0.1}, use_07_metric=True)
    _, _, oap7 = o.eval_det(pred_all, gt_all, {'bed': 0.25, 'chair': 0.5, 'table': 0.1}, use_07_metric=True)
    for c in oap7:
        assert abs(ap7[c] - oap7[c]) < 1e-12


def test_match_overlaps_and_edge_cases():
    from transferable3d_b200 import eval_det as ed
    from oracle import eval_det as o, box_util as ob
    box = lambda x: ob.get_3d_box((1.0, 1.0, 1.0), 0.0, (x, 0.0, 0.0))
    gt = {7: [box(0.0), box(10.0)], 9: []}
    pred = {7: [(box(0.0), 0.9), (box(0.05), 0.8), (box(10.5), 0.7), (box(50.0), 0.6)], 8: [(box(0.0), 0.95)]}
    m = ed.match_detections(pred, gt, 0.25)
    assert m['npos'] == 2
    assert m['tp'].tolist() == [0, 1, 0, 1, 0] and m['fp'].tolist() == [1, 0, 1, 0, 1]     # score order: .95 (image without GT), .9, .8, .7, .6
    assert abs(m['ovmax'][3] - 1.0 / 3.0) < 1e-5 and m['jmax'].tolist() == [-1, 0, 0, 1, 0]
    rec, prec, ap = ed.eval_det_cls(pred, gt, 0.25)
    orec, oprec, oap = o.eval_det_cls(pred, gt, 0.25)
    assert np.array_equal(rec, orec) and np.array_equal(prec, oprec) and abs(ap - oap) < 1e-12
    assert abs(ed.get_iou(box(0.0), box(0.5)) - 1.0 / 3.0) < 1e-5


def test_prediction_corners_vs_oracle():
    """evaluate.py:53-67: class2angle / class2size -> get_3d_box -> rotate_pc_along_y(-rot_angle)."""
    from transferable3d_b200 import eval_det as ed
    from oracle import box_util as ob, roi_seg_box3d_dataset as ods
    rng = np.random.RandomState(5)
    B = 64
    center, hc, sc = rng.randn(B, 3) * 2, rng.randint(0, 12, B), rng.randint(0, 10, B)
    hres, sres, rot = rng.randn(B) * 0.1, rng.randn(B, 3) * 0.1, rng.uniform(-np.pi, np.pi, B)
    got = ed.prediction_corners(center, hc, hres, sc, sres, rot)
    for i in range(B):
        ang = ods.class2angle(hc[i], hres[i], 12)
        want = ods.rotate_pc_along_y(ob.get_3d_box(ods.class2size(sc[i], sres[i]), ang, center[i]), -rot[i])
        assert np.abs(got[i] - want).max() < 1e-5
