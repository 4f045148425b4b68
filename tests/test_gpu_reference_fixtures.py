"""-m gpu: the CUDA path (through the C ABI) against the committed REFERENCE fixtures tests/golden/ref_*.npz -- outputs of the
reference's own unmodified source run on the TF1 stand-in (tests/golden/make_reference_golden.py; float64) -- on the same seeded
inputs, weights, flags and dropout masks, with no oracle in between:

  * the test graph of model F (test_semisup.get_model: refine = 2, masked point cloud, oracle mask, normalised 2D box feature)
    in the fp32 and f16x2 modes: mask logits within 1e-4 of the logit scale, every box output within 2e-4 of its scale;
  * one training step of BoxPC-Fit (representations A and B), of the semi-supervised adversarial graph (cfg5 flags) and of
    model A, at B = 8, N = 256: loss within 2e-4, head outputs within 2e-3, the set of trained variables, every gradient through
    the fixture's signature (norm within 2 %, four seeded random projections within 6 % of the norm: the whole-tensor bound of
    tests/util.py:assert_grad_close -- ReLU flips near zero move a weight gradient by ~1e-3), the updated moving statistics.
    The batches are drawn with seeds whose max-pooled columns have no near-tie (reference_cases.find_tie_free_seed) and whose
    smallest mask-logit margin is > 4e-4, so every assertion is unconditional.
  * the same training steps under every flag set the fixtures hold (8 BoxPC loss / weighting sets, 7 reprojection / intra-class /
    inactive-volume sets of the cfg5 graph, T-Net-only training, model A with the softmax projection): loss, trained-variable set,
    gradient signatures;
  * the numpy side on the device: ROISegBoxDataset batch assembly, eval_det, compute_box3d_iou.
"""
import os
import sys

import numpy as np
import pytest
import torch

from util import err_stats

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
sys.path.insert(0, GOLDEN)

if torch.cuda.is_available():
    import reference_cases as rc
    from transferable3d_b200 import runtime as rt, test_semisup as ts, config
    from transferable3d_b200 import train_boxpc as tb, train_semisup_adv as tsa, train_semisup as tsm

DEV = 'cuda:0'


def _fixture(name):
    return dict(np.load(os.path.join(GOLDEN, 'ref_%s.npz' % name)))


def _close(got, want, tol, what, floor=1e-6):
    got = np.asarray(got.detach().float().cpu().numpy() if isinstance(got, torch.Tensor) else got, dtype=np.float64).reshape(np.shape(want))
    assert np.isfinite(got).all(), what
    scale = max(float(np.abs(want).mean()), floor)
    err = float(np.abs(got - want).max())
    assert err <= tol * scale, (what, err, scale)


@pytest.mark.parametrize('mode', ['fp32', 'f16x2'])
@pytest.mark.parametrize('case,refine,mask_pc,oracle_mask,box2d', [('refine2', 2, False, False, False), ('masked_pc', 1, True, False, False),
                                                                   ('oracle_mask', 1, False, True, False), ('box2d_feats', 1, False, False, True),
                                                                   # an empty, a full and a one-point oracle mask; two refinement steps on the
                                                                   # masked cloud
                                                                   ('edge_masks', 2, True, True, False)])
def test_model_F_test_graph_vs_reference_fixture(case, refine, mask_pc, oracle_mask, box2d, mode, built_lib):
    want = _fixture('model_F_test_graph_' + case)
    v, b = rc._model_F_inputs(box2d_feats=box2d, edge_masks=(case == 'edge_masks'))
    FLAGS = config.cfg(refine=refine, mask_pc_for_boxpc=mask_pc, USE_NORMALIZED_BOX2D_AS_FEATS=box2d)
    rt.set_default_store(rt.VariableStore(v, DEV))
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32)).to(DEV)
    with rt.precision(mode), torch.no_grad():
        logits, ep = ts.build_graph(FLAGS, t(b['pc']), t(b['one_hot']), box2D=t(b['box2D']), img_dim=t(b['img_dim']),
                                    oracle_mask=t(b['labels']) if oracle_mask else None)
    torch.cuda.synchronize()
    # every end point the reference's graph exposes exists under the same key
    ref_keys = {k[3:].split('#')[0] for k in want if k.startswith('ep.')}
    ref_keys = {k[:-2] if k.startswith('F_pred_box_reg.') else k for k in ref_keys}
    assert not [k for k in sorted(ref_keys) if k not in ep], [k for k in sorted(ref_keys) if k not in ep]
    wl = want['logits']
    margin = np.abs(wl[..., 1] - wl[..., 0]).min()
    assert margin > 1e-2 * np.abs(wl).mean()                       # no point near the strict compare: the masks must be identical
    _close(logits, wl, 1e-4, 'logits')
    assert np.array_equal((logits[..., 0] < logits[..., 1]).cpu().numpy(), wl[..., 0] < wl[..., 1])
    for k in ('stage1_center', 'F_center', 'F_heading_scores', 'F_heading_residuals', 'F_size_scores', 'F_size_residuals', 'F2_center',
              'F2_heading_residuals', 'F2_size_residuals', 'boxpc_fit_prob'):
        _close(ep[k], want['ep.' + k], 2e-4, k)
    for k in ('center', 'heading_scores', 'heading_residuals', 'size_scores', 'size_residuals', 'box_params'):      # the W_ branch, where exposed
        if isinstance(ep.get(k), torch.Tensor):
            _close(ep[k], want['ep.' + k], 2e-4, k)
    for i, part in enumerate(ep['F_pred_box_reg']):
        _close(part, want['ep.F_pred_box_reg.%d' % i], 2e-4, 'F_pred_box_reg.%d' % i)


def _check_grads(want, grads, strip=''):
    """grads: {name: tensor} of the CUDA step.  Same trained variables, same None pattern, signatures inside the whole-tensor bound."""
    names = sorted(k[len('has_grad.'):] for k in want if k.startswith('has_grad.'))
    assert int(want['n_trained'][0]) == len(names)
    reached = [n for n in names if want['has_grad.' + n][0] > 0]
    assert sorted(grads) == sorted(n[len(strip):] for n in reached), sorted(set(grads) ^ set(n[len(strip):] for n in reached))
    wnorm = {}
    for n in reached:
        if n.endswith('weights'):
            wnorm[n.rsplit('/', 1)[0]] = float(want['grad.' + n + '#sig'][0])
    for n in reached:
        ref = want['grad.' + n + '#sig']
        got = rc.sig(grads[n[len(strip):]].detach().float().cpu().numpy())
        layer = n.rsplit('/', 1)[0] if not n.endswith(('gamma', 'beta')) else n.rsplit('/', 2)[0]
        floor = 1e-2 * wnorm.get(layer, 0.0) + 1e-7               # a bias in front of a batch norm has an analytically zero gradient
        assert np.isfinite(got).all(), n
        assert abs(got[0] - ref[0]) <= 2e-2 * ref[0] + floor, (n, got[0], ref[0])
        assert np.abs(got[2:] - ref[2:]).max() <= 6e-2 * ref[0] + 3 * floor, (n, got, ref)


def _check_moving(want, moving, prefix=''):
    for k, mv in moving.items():
        ref = want['moving.' + prefix + k]
        small = k.endswith('variance') and 'fc' in k.split('/')[-3]
        s = err_stats(mv.cpu().numpy().reshape(ref.shape), ref)
        assert s['max_abs'] <= (2e-2 if small else 5e-4) * max(s['ref_scale'], 1e-3), (k, s)


@pytest.mark.parametrize('rep', ['A', 'B'])
def test_boxpc_train_step_vs_reference_fixture(rep, built_lib):
    want = _fixture('boxpc_train_rep_' + rep)
    v, feed, masks = rc._boxpc_inputs(rep)
    B, N = feed['pc'].shape[:2]
    FLAGS = config.cfg(BOXPC_WEIGHT_DELTA=4., BOX_PC_MASK_REPRESENTATION=rep)
    g = tb.BoxPCTrainGraph(v, FLAGS, B, N, 6, DEV)
    out = g.forward_backward(feed, masks)
    torch.cuda.synchronize()
    assert abs(float(out['loss']) - want['loss'][0]) <= 2e-4 * max(1.0, abs(want['loss'][0]))
    assert np.abs(out['boxpc_delta_center'].cpu().numpy() - want['ep.boxpc_delta_center']).max() <= 2e-4
    for k in ('boxpc_delta_size', 'boxpc_delta_angle', 'boxpc_fit_logits'):
        if isinstance(out.get(k), torch.Tensor):
            assert np.abs(out[k].cpu().numpy().reshape(want['ep.' + k].shape) - want['ep.' + k]).max() <= 2e-4, k
    _check_grads(want, g.grad, strip='box_pc_mask_model/')
    _check_moving(want, g.moving, prefix='box_pc_mask_model/')
    from oracle.train_boxpc import get_learning_rate, get_bn_decay          # schedules (plain arithmetic) against the script's own
    assert abs(get_learning_rate(0, B) - want['learning_rate'][0]) < 1e-12 and abs(get_bn_decay(0, B) - want['bn_decay'][0]) < 1e-12


def test_semisup_adv_train_step_vs_reference_fixture(built_lib):
    want = _fixture('semisup_adv_train_cfg5')
    v, feed, masks = rc._semi_inputs('F')
    B, N = feed['pc'].shape[:2]
    g = tsa.SemiAdvTrainGraph(v, config.cfg(**rc.CFG5), B, N, 6, DEV)
    ep = g.forward_backward(feed, masks)
    torch.cuda.synchronize()
    loss = float(ep['loss_terms'].cpu().numpy()[0])
    assert abs(loss - want['loss'][0]) <= 2e-4 * max(1.0, abs(want['loss'][0])), (loss, want['loss'][0])
    for k in ('stage1_center', 'F_center', 'F_heading_scores', 'F_size_residuals', 'boxpc_fit_prob', 'F2_center', 'F2_heading_residuals',
              'F2_size_residuals'):
        _close(ep[k], want['ep.' + k], 2e-3, k, floor=1.0)
    _check_grads(want, g.grad)
    _check_moving(want, g.moving)


def test_semisup_model_a_train_step_vs_reference_fixture(built_lib):
    want = _fixture('semisup_A_train')
    v, feed, masks = rc._semi_inputs('A')
    B, N = feed['pc'].shape[:2]
    g = tsm.SemiTrainGraph(v, config.cfg(SEMI_MODEL='A', **rc.CFG_A), B, N, 6, DEV)
    ep = g.forward_backward(feed, masks)
    torch.cuda.synchronize()
    assert abs(float(ep['semi_loss']) - want['loss'][0]) <= 2e-4 * max(1.0, abs(want['loss'][0])), (float(ep['semi_loss']), want['loss'][0])
    for k in ('stage1_center', 'center', 'heading_scores', 'size_residuals'):
        _close(ep[k], want['ep.' + k], 2e-3, k, floor=1.0)
    _check_grads(want, g.grad)
    _check_moving(want, g.moving)


# ---- the numpy side of the reference, device counterparts -----------------------------------------------------------------------
def test_dataset_get_batch_vs_reference_fixture(tmp_path, built_lib):
    """transferable3d_b200.roi_seg_box3d_dataset.ROISegBoxDataset (pickle read on the host, batch assembly in one kernel) against what
    the reference's own class returned for the same gzip-pickled file, the same index permutation and the same numpy seed."""
    import gzip
    import pickle
    from transferable3d_b200 import roi_seg_box3d_dataset as D
    want = _fixture('dataset_get_batch')
    path = os.path.join(str(tmp_path), 'frustums.zip.pickle')
    with gzip.open(path, 'wb') as f:
        pickle.dump(rc._frustum_lists(), f, 2)
    names = ('data', 'image', 'label', 'center', 'hcls', 'hres', 'scls', 'sres', 'box2d', 'rtilts', 'ks', 'rot_angle', 'img_dims', 'one_hot')
    tol = dict(data=2e-5, center=2e-5, hres=1e-5, sres=1e-6, box2d=1e-4, rtilts=1e-6, ks=1e-4, rot_angle=1e-6, img_dims=1e-4)
    for i, (rotate, flip, shift, one_hot) in enumerate(rc.DATASET_VARIANTS):
        ds = D.ROISegBoxDataset(rc.DATASET_CLASSES, 256, 'val', random_flip=flip, random_shift=shift, rotate_to_center=rotate,
                                overwritten_data_path=path, one_hot=one_hot)
        assert len(ds) == int(want['v%d.len' % i][0])
        idxs = np.random.RandomState(1).permutation(len(ds))
        np.random.seed(77)
        got = ds.get_batch(idxs, 4, 20, 256, 6)
        assert len(got) == (14 if one_hot else 13) and got[1] is None
        for name, g in zip(names, got):
            if name == 'image':
                continue
            g = g.cpu().numpy().astype(np.float64)
            key = 'v%d.%s' % (i, name)
            if key in want:
                w = want[key]
                if name in ('label', 'hcls', 'scls', 'one_hot'):
                    assert np.array_equal(g.reshape(w.shape), w), key
                else:
                    assert np.abs(g.reshape(w.shape) - w).max() <= tol[name] * max(1.0, np.abs(w).max()), key
            else:                                       # stored as a signature (the point tensor and the labels)
                w = want[key + '#sig']
                s = rc.sig(g)
                if name == 'label':
                    assert np.abs(s - w).max() <= 1e-9 * max(1.0, w[0]), key
                else:
                    assert abs(s[0] - w[0]) <= 1e-5 * w[0] and np.abs(s[2:] - w[2:]).max() <= 1e-4 * w[0], (key, s, w)


def test_eval_det_and_box_iou_vs_reference_fixture(built_lib):
    """eval_det (matching loop on the device) and compute_box3d_iou against the reference's eval_det.py / roi_seg_box3d_dataset.py run
    on the same detections; 3D IoU itself comes from the restated box_util in both (the reference tree lacks that file)."""
    from transferable3d_b200 import eval_det as E, box_util as gbu
    want = _fixture('numpy_helpers')
    pred_all, gt_all = rc._scene(0)
    for tag, thr, m07 in (('a', 0.25, False), ('b', {'bed': 0.25, 'chair': 0.5, 'table': 0.1}, True)):
        rec, prec, ap = E.eval_det(pred_all, gt_all, thr, use_07_metric=m07)
        assert sorted(ap) == ['bed', 'chair', 'table']
        for c in ap:
            assert np.array_equal(np.asarray(rec[c], dtype=np.float64), want['eval_det.%s.%s.rec' % (tag, c)]), (tag, c)
            assert np.array_equal(np.asarray(prec[c], dtype=np.float64), want['eval_det.%s.%s.prec' % (tag, c)]), (tag, c)
            assert abs(ap[c] - want['eval_det.%s.%s.ap' % (tag, c)][0]) < 1e-12
    # the same draws as reference_cases._numpy_side makes before compute_box3d_iou
    side = rc._numpy_side(_Recorder(), _Recorder(), None, None, host_scalars_only=False)
    args = side['_compute_box3d_iou_args']
    i2, i3 = gbu.compute_box3d_iou(*args)
    got = np.stack([i2.cpu().numpy(), i3.cpu().numpy()]).astype(np.float64)
    assert np.abs(got - want['compute_box3d_iou']).max() < 2e-5
    assert (want['compute_box3d_iou'] > 0.01).any()


class _Recorder(object):
    """Stands in for a module in reference_cases._numpy_side: every function accepts its arguments and returns a placeholder, so the
    random draws up to compute_box3d_iou are consumed in the same order and its arguments come back."""
    def __getattr__(self, name):
        if name in ('angle2class', 'size2class'):
            return lambda *a, **k: (0, np.zeros(3) if name == 'size2class' else 0.0)
        if name == 'from_prediction_to_label_format':
            return lambda *a, **k: (0.0,) * 7
        if name == 'rotate_pc_along_y':
            return lambda pc, a: pc
        if name == 'get_3d_box':
            return lambda *a, **k: np.zeros((8, 3))
        return lambda *a, **k: 0.0


# ---- every flag set of the training graphs (loss, trained-variable set and gradient signatures; fixtures keep nothing else) --------
@pytest.mark.parametrize('i', range(8))
def test_boxpc_flag_variants_vs_reference_fixture(i, built_lib):
    want = _fixture('boxpc_train_variant%d' % i)
    rep = 'A' if i != 7 else 'B'
    v, feed, masks = rc._boxpc_inputs(rep)
    B, N = feed['pc'].shape[:2]
    g = tb.BoxPCTrainGraph(v, config.cfg(BOXPC_WEIGHT_DELTA=4., BOX_PC_MASK_REPRESENTATION=rep, **rc.BOXPC_VARIANTS[i]), B, N, 6, DEV)
    out = g.forward_backward(feed, masks)
    torch.cuda.synchronize()
    assert abs(float(out['loss']) - want['loss'][0]) <= 2e-4 * max(1.0, abs(want['loss'][0])), (float(out['loss']), want['loss'][0])
    _check_grads(want, g.grad, strip='box_pc_mask_model/')


@pytest.mark.parametrize('i', range(7))          # variant 7 (two refinement steps) is refused by the training graph: see below
def test_semisup_adv_flag_variants_vs_reference_fixture(i, built_lib):
    want = _fixture('semisup_adv_train_variant%d' % i)
    v, feed, masks = rc._semi_inputs('F')
    B, N = feed['pc'].shape[:2]
    g = tsa.SemiAdvTrainGraph(v, config.cfg(**dict(rc.CFG5, **rc.REPROJ_VARIANTS[i])), B, N, 6, DEV)
    ep = g.forward_backward(feed, masks)
    torch.cuda.synchronize()
    loss = float(ep['loss_terms'].cpu().numpy()[0])
    assert abs(loss - want['loss'][0]) <= 2e-4 * max(1.0, abs(want['loss'][0])), (loss, want['loss'][0])
    _check_grads(want, g.grad)


def test_semisup_adv_variants_outside_the_recipe_fail_loudly(built_lib):
    v, feed, masks = rc._semi_inputs('F')
    with pytest.raises(NotImplementedError):
        tsa.SemiAdvTrainGraph(v, config.cfg(**dict(rc.CFG5, **rc.REPROJ_VARIANTS[7])), 8, 256, 6, DEV)


def test_semisup_adv_tnet_only_and_model_a_softmax_vs_reference_fixture(built_lib):
    want = _fixture('semisup_adv_train_tnet_only')
    v, feed, masks = rc._semi_inputs('F')
    g = tsa.SemiAdvTrainGraph(v, config.cfg(**dict(rc.CFG5, SEMI_TRAIN_BOX_TRAIN_CLASS_AG_BOX=False)), 8, 256, 6, DEV)
    ep = g.forward_backward(feed, masks)
    torch.cuda.synchronize()
    assert abs(float(ep['loss_terms'].cpu().numpy()[0]) - want['loss'][0]) <= 2e-4 * max(1.0, abs(want['loss'][0]))
    _check_grads(want, g.grad)
    want = _fixture('semisup_A_train_softmax_proj')
    v, feed, masks = rc._semi_inputs('A')
    flags = dict(rc.CFG_A, WEAK_TRAIN_BOX_W_SURFACE=[True, False, True], WEAK_REPROJECTION_USE_SOFTMAX_PROJ=True)
    g = tsm.SemiTrainGraph(v, config.cfg(SEMI_MODEL='A', **flags), 8, 256, 6, DEV)
    ep = g.forward_backward(feed, masks)
    torch.cuda.synchronize()
    assert abs(float(ep['semi_loss']) - want['loss'][0]) <= 2e-4 * max(1.0, abs(want['loss'][0]))
    _check_grads(want, g.grad)


def test_fpointnet_v1_helpers_vs_reference_fixture(built_lib):
    """models/model_util.py on the device against the reference's own functions: mask + centroid, the 512-point gather on numpy's
    legacy stream (same seed: > 512, a handful and no selected points), the T-Net on the gathered points, the KITTI-sized output
    parse, the corner builders with the double residual add."""
    from transferable3d_b200 import model_util as mu
    want = _fixture('fpointnet_v1_helpers')
    v, b, logits, lab, out59 = rc._fpn_inputs()
    rt.set_default_store(rt.VariableStore(v, DEV))
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32)).to(DEV)

    def sig_close(got, key, rel):
        w = want[key + '#sig']
        s = rc.sig(got.detach().float().cpu().numpy())
        assert abs(s[0] - w[0]) <= rel * max(w[0], 1e-6) and np.abs(s[2:] - w[2:]).max() <= 10 * rel * max(w[0], 1e-6), (key, s, w)
    try:
        with rt.precision('fp32'), torch.no_grad():
            ep = {}
            mu.set_resample_rng('numpy_legacy', seed=1234)
            obj, mean, ep = mu.point_cloud_masking(t(b['pc']), t(logits), ep)
            with rt.variable_scope('tnet'):
                delta, ep = mu.get_center_regression_net(obj, t(b['one_hot']), False, None, ep)
            mu.set_resample_rng('numpy_legacy', seed=99)
            obj6, _, _ = mu.point_cloud_masking(t(b['pc']), t(logits), {}, xyz_only=False)
            ep = mu.parse_output_to_tensors(t(out59), ep)
            center = ep['center_boxnet'] + delta + mean
            sun_res = t((np.random.RandomState(3).standard_normal((4, 10, 3)) * 0.1).astype(np.float32))
            ck = mu.get_box3d_corners(center, ep['heading_residuals'], ep['size_residuals'])
            cs = mu.get_box3d_corners_sunrgbd(center, ep['heading_residuals'], sun_res)
            ch = mu.get_box3d_corners_helper(t(lab['center']), t(lab['hres']), t(np.abs(lab['sres']) + 0.5))
        torch.cuda.synchronize()
    finally:
        mu.set_resample_rng('philox', 0)
    assert tuple(obj.shape) == (4, 512, 3) and tuple(obj6.shape) == (4, 512, 6)
    sig_close(obj, 'object_pc', 1e-5)                                   # the same 512 picks per frustum, in the same order
    sig_close(obj6, 'object_pc_6ch', 1e-5)
    sig_close(ep['mask'], 'ep.mask', 1e-9)
    _close(mean, want['mask_xyz_mean'], 1e-5, 'mask_xyz_mean', floor=1.0)
    _close(delta, want['tnet_delta'], 2e-4, 'tnet_delta', floor=1e-2)
    for k in ('center_boxnet', 'heading_scores', 'heading_residuals_normalized', 'heading_residuals', 'size_scores',
              'size_residuals_normalized', 'size_residuals'):
        _close(ep[k], want['ep.' + k], 1e-5, k, floor=1e-2)
    sig_close(ck, 'corners_kitti', 2e-5)
    sig_close(cs, 'corners_sunrgbd', 2e-5)
    _close(ch, want['corners_helper'], 1e-5, 'corners_helper', floor=1.0)
    from transferable3d_b200.constants import g_mean_size_arr, MEAN_DIMS_ARR
    assert np.array_equal(np.asarray(g_mean_size_arr, dtype=np.float64), want['g_mean_size_arr'])
    assert np.array_equal(np.asarray(MEAN_DIMS_ARR, dtype=np.float64), want['sun_mean_size_arr'])


@pytest.mark.parametrize('mode', ['fp32', 'f16x2'])
def test_model_F_test_graph_2048_points_vs_reference_fixture(mode, built_lib):
    """The test graph at the reference's own point count (8 frustums x 2048 points: 16 tiles per frustum through the fused
    chains) against what the reference's source computes; the 32768 logits are compared through their signature."""
    want = _fixture('model_F_test_graph_2048_points')
    v, b = rc._model_F_inputs(full_size=True)
    rt.set_default_store(rt.VariableStore(v, DEV))
    t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32)).to(DEV)
    with rt.precision(mode), torch.no_grad():
        logits, ep = ts.build_graph(config.cfg(refine=1), t(b['pc']), t(b['one_hot']), box2D=t(b['box2D']), img_dim=t(b['img_dim']))
    torch.cuda.synchronize()
    assert tuple(logits.shape) == (8, 2048, 2)
    w = want['logits#sig']
    s = rc.sig(logits.float().cpu().numpy())
    assert abs(s[0] - w[0]) <= 1e-5 * w[0] and np.abs(s[2:] - w[2:]).max() <= 1e-4 * w[0], (s, w)
    # every reference margin is > 0.04 (9 % of the logit scale; reference_cases / make_reference_golden print it): all points masked in
    assert bool((logits[..., 0] < logits[..., 1]).all())
    for k in ('stage1_center', 'F_center', 'F_heading_scores', 'F_heading_residuals', 'F_size_scores', 'F_size_residuals', 'F2_center',
              'F2_heading_residuals', 'F2_size_residuals', 'boxpc_fit_prob', 'box_params'):
        _close(ep[k], want['ep.' + k], 2e-4, k)
    for key in ('feats_lv1', 'feats_lv2', 'feats_lv3'):
        w = want['ep.%s#sig' % key]
        s = rc.sig(ep[key].float().cpu().numpy())
        assert abs(s[0] - w[0]) <= 1e-4 * w[0] and np.abs(s[2:] - w[2:]).max() <= 1e-3 * w[0], (key, s, w)
