"""-m gpu: frustum batch assembly (input side): ROISegBoxDataset.get_batch on the device
(t3d_assemble_frustum_batch) against the literal numpy restatement of __getitem__ + get_batch under the same numpy seed,
through a gz-pickle file in the reference's 13-list (and 7-list) layout."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CLASSES = ['bed', 'table', 'sofa', 'chair', 'toilet', 'desk', 'dresser', 'night_stand', 'bookshelf', 'bathtub']


def _lists(F, seed, with_extra_class=True):
    from oracle import box_util as ob
    from transferable3d_b200.constants import type_mean_size
    rng = np.random.RandomState(seed)
    L = [[] for _ in range(13)]
    for i in range(F):
        cls = CLASSES[rng.randint(10)] if not (with_extra_class and i % 9 == 4) else 'lamp'      # filtered out by `classes`
        n = rng.randint(40, 700)
        size = type_mean_size.get(cls, np.ones(3)) * rng.uniform(0.8, 1.2, 3)
        heading = rng.uniform(-np.pi, np.pi) + 0.013                      # keep clear of the angle-bin boundaries
        center = np.array([rng.uniform(-2, 2), rng.uniform(-0.5, 0.5), rng.uniform(1.5, 5)])
        pts = np.concatenate([center + rng.randn(n, 3) * 0.5, rng.rand(n, 3)], axis=1)
        rec = [1000 + i, rng.rand(4) * 400, ob.get_3d_box(size, heading, center), None, pts, rng.randint(0, 2, n).astype(np.int64), cls,
               heading, size, np.eye(3) + rng.randn(3, 3) * 0.01, np.diag([520.0, 520.0, 1.0]) + rng.rand(3, 3), rng.uniform(-0.6, 0.6),
               np.array([640.0, 480.0])]
        for l, v in zip(L, rec):
            l.append(v)
    return L


@pytest.mark.parametrize('rotate,flip,shift,one_hot', [(True, False, False, True), (False, False, False, False), (True, True, True, True)])
def test_get_batch_vs_oracle(tmp_path, rotate, flip, shift, one_hot):
    from transferable3d_b200 import roi_seg_box3d_dataset as D, utils
    from oracle import roi_seg_box3d_dataset as OD
    lists = _lists(60, seed=3)
    path = os.path.join(str(tmp_path), 'frustums.zip.pickle')
    utils.save_zipped_pickle(lists, path)
    ds = D.ROISegBoxDataset(CLASSES, 512, 'val', random_flip=flip, random_shift=shift, rotate_to_center=rotate,
                            overwritten_data_path=path, one_hot=one_hot)
    keep = [i for i, c in enumerate(lists[6]) if c in CLASSES]
    assert len(ds) == len(keep) < 60 and ds.idx_l == [lists[0][i] for i in keep]
    ods = OD.ROISegBoxDataset([[l[i] for i in keep] for l in lists], 512, random_flip=flip, random_shift=shift,
                              rotate_to_center=rotate, one_hot=one_hot)
    idxs = np.random.RandomState(1).permutation(len(ds))
    np.random.seed(77)
    got = ds.get_batch(idxs, 4, 36, 512, 6)
    np.random.seed(77)
    want = ods.get_batch(idxs, 4, 36, 512, 6)
    assert len(got) == len(want) == (14 if one_hot else 13) and got[1] is None
    G = lambda t: t.cpu().numpy()
    assert np.abs(G(got[0]) - want[0]).max() < 2e-5                       # fp32 rotation vs float64 numpy, coordinates up to ~6 m
    if not (rotate or flip or shift):
        assert np.array_equal(G(got[0]), want[0].astype(np.float32))      # pure gather: bit-exact
    assert np.array_equal(G(got[2]), want[2])                             # seg labels
    assert np.abs(G(got[3]) - want[3]).max() < 2e-6 * 6 + 1e-6            # box centre
    assert np.array_equal(G(got[4]), want[4]) and np.abs(G(got[5]) - want[5]).max() < 1e-5      # heading class / residual
    assert np.array_equal(G(got[6]), want[6]) and np.abs(G(got[7]) - want[7]).max() < 1e-6      # size class / residual
    for k in (8, 9, 10, 12):
        assert np.allclose(G(got[k]), want[k], rtol=1e-6, atol=1e-6)
    assert np.abs(G(got[11]) - want[11]).max() < 1e-6                     # rot_angle
    if one_hot:
        assert np.array_equal(G(got[13]), want[13])


def test_rgb_detection_layout_and_pickle_roundtrip(tmp_path):
    from transferable3d_b200 import roi_seg_box3d_dataset as D, utils
    L13 = _lists(20, seed=5, with_extra_class=False)
    rng = np.random.RandomState(0)
    L7 = [L13[0], L13[1], L13[3], L13[4], L13[6], L13[11], list(rng.rand(20))]
    path = os.path.join(str(tmp_path), 'rgb_det.zip.pickle')
    utils.save_zipped_pickle(L7, path)
    back = utils.load_zipped_pickle(path)
    assert len(back) == 7 and back[4] == L7[4] and np.array_equal(back[3][7], L7[3][7])
    ds = D.ROISegBoxDataset(CLASSES, 256, 'val', rotate_to_center=True, overwritten_data_path=path, from_rgb_detection=True, one_hot=True)
    np.random.seed(5)
    data, img, rot, prob, one_hot, y_seg = ds.get_batch(np.arange(20), 0, 8, 256, 6, from_rgb_detection=True)
    np.random.seed(5)
    for i in range(8):
        choice = np.random.choice(L7[3][i].shape[0], 256, replace=True)
        want = D.rotate_pc_along_y(np.copy(L7[3][i]), np.pi / 2 + L7[5][i])[choice]
        assert np.abs(data[i].cpu().numpy() - want).max() < 2e-5
        assert abs(float(rot[i]) - (np.pi / 2 + L7[5][i])) < 1e-6 and abs(float(prob[i]) - L7[6][i]) < 1e-6
    assert img is None and one_hot.shape == (8, 10) and float(y_seg.abs().max()) == 0.0
    with pytest.raises(AssertionError):
        D.ROISegBoxDataset(CLASSES, 256, 'val', overwritten_data_path=path)           # 7 lists where 13 are expected


def test_main_batch_end_to_end(tmp_path):
    """Frustum file -> ROISegBoxDataset.get_batch -> model F + BoxPC refine (sess.run) -> device post-processing ->
    predictions / result files -> 3D AP: the reference's test flow (test_semisup.main_batch, evaluate_predictions) end to
    end on the device pieces.  Checked against the literal host inference() on the same batches."""
    from transferable3d_b200 import roi_seg_box3d_dataset as D, utils, test_semisup as ts, eval_det as ed, weights, config, runtime as rt
    from transferable3d_b200.constants import class2type
    lists = _lists(45, seed=9, with_extra_class=False)
    path = os.path.join(str(tmp_path), 'val.zip.pickle')
    utils.save_zipped_pickle(lists, path)
    ds = D.ROISegBoxDataset(CLASSES, 1024, 'val', rotate_to_center=True, overwritten_data_path=path, one_hot=True)
    variables, _ = weights.standard_model_F()
    FLAGS = config.cfg()
    out_pickle, result_dir = os.path.join(str(tmp_path), 'pred.zip.pickle'), os.path.join(str(tmp_path), 'results')
    with rt.precision('fp32'):
        sess_ops = ts.get_model(32, 1024, 6, FLAGS=FLAGS, variables=variables, cuda_graph=False)
        np.random.seed(3)
        pred = ts.main_batch(ds, CLASSES, 10, 1024, 6, prefix='F2_', use_boxpc_fit_prob=True, sess_ops=sess_ops,
                             output_filename=out_pickle, result_dir=result_dir)
        # the same batches through the literal host post-processing
        np.random.seed(3)
        sess, ops = sess_ops
        ref_scores, ref_seg = [], []
        for s in range(0, 45, 32):
            e = min(45, s + 32)
            b = ds.get_batch(np.arange(45), s, e, 1024, 6)
            r = ts.inference(sess, ops, ts._pad_batch(b[0], 32).cpu().numpy(), ts._pad_batch(b[13], 32).cpu().numpy(), 32, prefix='F2_',
                             use_boxpc_fit_prob=True)
            ref_scores += list(r[6][:e - s])
            ref_seg += list(r[0][:e - s])
    assert len(pred) == 14 and all(len(l) == 45 for l in pred)
    assert np.allclose(np.array(pred[9]), np.array(ref_scores), rtol=1e-5, atol=1e-5)
    assert all(np.array_equal(a, b) for a, b in zip(pred[2], ref_seg))
    assert pred[11] == ds.idx_l and pred[0][0].shape == (1024, 6)
    back = utils.load_zipped_pickle(out_pickle)
    assert len(back) == 14 and np.array_equal(back[3][7], pred[3][7])
    n_lines = sum(len(open(os.path.join(result_dir, c + '_pred.txt')).read().splitlines()) for c in CLASSES)
    assert n_lines == 45
    # predictions -> boxes -> AP (evaluate.py:53-72); with random weights the AP is just a number in [0, 1]
    corners = ed.prediction_corners(pred[3], pred[4], pred[5], pred[6], pred[7], pred[8])
    pred_all, gt_all = {}, {}
    for i in range(45):
        pred_all.setdefault(pred[11][i], []).append((class2type[pred[10][i]], corners[i], pred[9][i]))
        gt_all.setdefault(ds.idx_l[i], []).append((ds.cls_type_l[i], ds.box3d_l[i]))
    rec, prec, ap = ed.eval_det(pred_all, gt_all, 0.25)
    assert set(ap) == set(ds.cls_type_l) and all(0.0 <= v <= 1.0 for v in ap.values())


def test_checkpoint_to_session(tmp_path):
    """tf_checkpoint.load_checkpoint -> get_model / sess.run: the loaded dict drives the model exactly like the dict it was
    written from (the path saver.restore takes in the reference, test_semisup.py:158-159)."""
    from transferable3d_b200 import tf_checkpoint as ck, test_semisup as ts, weights, synth, config, runtime as rt
    variables, _ = weights.standard_model_F()
    prefix = os.path.join(str(tmp_path), 'model.ckpt')
    ck.save_checkpoint(prefix, {k: np.asarray(v) for k, v in variables.items()})
    loaded = ck.load_checkpoint(prefix)
    b = synth.make_batch(4, 1024, 6, seed=2)
    FLAGS = config.cfg()
    outs = []
    for v in (variables, loaded):
        with rt.precision('fp32'):
            sess, ops = ts.get_model(4, 1024, 6, FLAGS=FLAGS, variables=v, cuda_graph=False)
            outs.append(sess.run([ops['logits'], ops['end_points']['F2_center']],
                                 {ops['pc_pl']: b['pc'], ops['one_hot_vec_pl']: b['one_hot'], ops['is_training_pl']: False}))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    # the reference's own calling convention: flags and checkpoint path in module globals, get_model(B, N, C) positional
    # (test_semisup.py:22, :61, :158-159, :545-548)
    with pytest.raises(ValueError):
        ts.get_model(4, 1024, 6)
    ts.FLAGS, ts.MODEL_PATH = FLAGS, prefix
    try:
        with rt.precision('fp32'):
            sess, ops = ts.get_model(4, 1024, 6, False, cuda_graph=False)
            out = sess.run([ops['logits'], ops['end_points']['F2_center']],
                           {ops['pc_pl']: b['pc'], ops['one_hot_vec_pl']: b['one_hot'], ops['is_training_pl']: False})
    finally:
        ts.FLAGS = ts.MODEL_PATH = None
    assert torch.equal(out[0], outs[0][0]) and torch.equal(out[1], outs[0][1])
