"""-m gpu: inference post-processing through the C ABI against the numpy float64 oracle:
t3d_inference_scores (test_semisup.py:236-258), t3d_prediction_to_label (roi_seg_box3d_dataset.py:461-466), the
device-side inference() and the result writers."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fetches(B, N, seed, ties=True):
    rng = np.random.RandomState(seed)
    logits = rng.randn(B, N, 2).astype(np.float32) * 3.0
    if ties:
        logits[:, ::7, 1] = logits[:, ::7, 0]          # exact ties -> class 0 (np.argmax takes the first maximum)
    logits[0, :, 1] = logits[0, :, 0] - 1.0           # empty mask
    logits[1, :, 1] = logits[1, :, 0] + 1.0           # full mask
    hs = rng.randn(B, 12).astype(np.float32) * 2.0
    hs[2, 3] = hs[2, 8] = hs[2].max() + 1.0           # tied maximum -> first index
    ss = rng.randn(B, 10).astype(np.float32) * 2.0
    hr = rng.randn(B, 12).astype(np.float32) * 0.2
    sr = rng.randn(B, 10, 3).astype(np.float32) * 0.3
    fp = rng.rand(B).astype(np.float32)
    return logits, hs, hr, ss, sr, fp


@pytest.mark.parametrize('B,N,use_fit', [(8, 2048, True), (5, 1000, False), (3, 37, True)])
def test_inference_scores_vs_oracle(B, N, use_fit):
    from transferable3d_b200 import test_semisup as ts
    from oracle import test_semisup as o
    logits, hs, hr, ss, sr, fp = _fetches(B, N, seed=B + N)
    D = lambda a: torch.as_tensor(a).cuda()
    out = ts.inference_scores(D(logits), D(hs), D(hr), D(ss), D(sr), D(fp) if use_fit else None)
    seg, mmp, hc, hres, sc, sres, score = o.inference_scores(logits, hs, hr, ss, sr, fp if use_fit else None)
    assert np.array_equal(out['pred_seg'].cpu().numpy(), seg)                       # bit-exact
    assert np.array_equal(out['heading_cls'].cpu().numpy(), hc) and np.array_equal(out['size_cls'].cpu().numpy(), sc)
    assert np.array_equal(out['heading_res'].cpu().numpy(), hres.astype(np.float32))
    assert np.array_equal(out['size_res'].cpu().numpy(), sres.astype(np.float32))
    assert np.allclose(out['mask_mean_prob'].cpu().numpy(), mmp, rtol=1e-5, atol=1e-6)
    assert np.allclose(out['scores'].cpu().numpy(), score, rtol=1e-5, atol=1e-5)      # fp32 sums / logs vs float64


def test_prediction_to_label_vs_oracle(tmp_path):
    from transferable3d_b200 import roi_seg_box3d_dataset as ds, test_semisup as ts
    from transferable3d_b200.constants import class2type
    from oracle import roi_seg_box3d_dataset as o
    rng = np.random.RandomState(3)
    B = 257
    center = rng.randn(B, 3) * 2.0
    hc, sc = rng.randint(0, 12, B), rng.randint(0, 10, B)
    hres, sres = rng.randn(B) * 0.2, rng.randn(B, 3) * 0.2
    rot = rng.uniform(-np.pi, np.pi, B)
    got = ds.from_prediction_to_label_format_batch(center, hc, hres, sc, sres, rot).cpu().numpy()
    want = np.array([o.from_prediction_to_label_format(center[i], hc[i], hres[i], sc[i], sres[i], rot[i]) for i in range(B)])
    # angles near the +pi wrap can differ by 2 pi between fp32 and float64 evaluation of `angle > pi`: compare mod 2 pi
    d = got - want
    d[:, 6] = (d[:, 6] + np.pi) % (2 * np.pi) - np.pi
    assert np.abs(d).max() < 2e-5, np.abs(d).max(0)
    one = ds.from_prediction_to_label_format(center[5], hc[5], hres[5], sc[5], sres[5], rot[5])
    assert np.allclose(one, got[5], atol=1e-6)
    # writers: same text format as the reference (test_semisup.py:262-294)
    types = [class2type[int(c)] for c in sc]
    box2d = rng.rand(B, 4) * 100
    scores = rng.randn(B)
    classes = sorted(set(class2type.values()))
    ts.write_detection_results(str(tmp_path), classes, list(range(B)), types, box2d, center, hc, hres, sc, sres, rot, scores)
    lines = []
    for c in classes:
        lines += open(os.path.join(str(tmp_path), c + '_pred.txt')).read().splitlines()
    assert len(lines) == B
    rec = {int(l.split()[0]): l.split() for l in lines}
    for i in (0, 100, 256):
        f = rec[i]
        assert f[1] == types[i] and f[2:5] == ['-1', '-1', '-10'] and len(f) == 17
        assert np.allclose([float(v) for v in f[5:9]], box2d[i], atol=1e-5)
        assert np.allclose([float(v) for v in f[9:15]], got[i, :6], atol=1e-5)
        assert abs(float(f[16]) - scores[i]) < 1e-5
    ts.fill_files(str(tmp_path), ['x_pred.txt'])
    assert os.path.exists(os.path.join(str(tmp_path), 'x_pred.txt'))


def test_inference_device_matches_host_inference():
    """The device post-processing returns the 7-tuple of the reference-literal inference() on the same session."""
    from transferable3d_b200 import test_semisup as ts, weights, synth, config, runtime as rt
    variables, _ = weights.standard_model_F()
    b = synth.make_batch(8, 2048, 6, seed=21)
    FLAGS = config.cfg()
    with rt.precision('fp32'):
        sess, ops = ts.get_model(4, 2048, 6, FLAGS=FLAGS, variables=variables, cuda_graph=False)
        ref = ts.inference(sess, ops, b['pc'], b['one_hot'], 4, prefix='F2_', use_boxpc_fit_prob=True)
        got = ts.inference_device(sess, ops, b['pc'], b['one_hot'], 4, prefix='F2_', use_boxpc_fit_prob=True)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[2], ref[2]) and np.array_equal(got[4], ref[4])
    for k in (1, 3, 5):
        assert np.allclose(got[k], ref[k], rtol=0, atol=1e-6)
    assert np.allclose(got[6], ref[6], rtol=1e-5, atol=1e-5)
