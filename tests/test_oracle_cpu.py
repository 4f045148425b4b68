"""CPU tests of the oracle itself: the only numeric example the reference holds (softmax weight
table, models/config.py:137-142), published Philox KAT vectors, numpy-stream equivalences behind
mask_to_indices, closed forms checked against literal restatements, tiny hand-computed cases,
and the committed golden fixtures."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import tf_util, model_util, semisup_models, weak_losses
from oracle.tf_layers import VarStore
from transferable3d_b200 import weights, synth, config
from transferable3d_b200.constants import MEAN_DIMS_ARR

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


def test_softmax_weight_table_from_reference_config():
    # models/config.py:137-142 -- the reference's only numeric example
    pts = np.array([0, 0.1, 0.2, 0.4, 0.6, 0.8, 0.9, 1.0])
    table = {1: [0.071, 0.079, 0.087, 0.106, 0.13, 0.158, 0.175, 0.194],
             5: [0.003, 0.005, 0.008, 0.023, 0.062, 0.168, 0.276, 0.455],
             10: [0.0, 0.0, 0.0, 0.002, 0.012, 0.089, 0.241, 0.656],
             20: [0.0, 0.0, 0.0, 0.0, 0.0, 0.016, 0.117, 0.867],
             40: [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.018, 0.982]}
    p2 = torch.tensor(np.stack([pts, pts], axis=1))[None]               # (1,8,2)
    for scale, w in table.items():
        # right-closeness softmax weights are what tf_get_2D_softmax_bbox_of_points uses for `right`
        got = torch.softmax(torch.tensor((pts - pts.min()) / (pts.max() - pts.min()) * scale), dim=0).numpy()
        assert np.allclose(got, w, atol=6e-4), (scale, got)
        box = tf_util.tf_get_2D_softmax_bbox_of_points(p2, float(scale))[0].numpy()
        assert abs(box[2] - float((pts * got).sum())) < 1e-9
        assert box[0] < box[2] and abs(box[0] - (1 - box[2])) < 1e-9      # symmetric point set


def test_philox_published_kat_vectors():
    # Random123 kat_vectors, philox4x32 10 rounds
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kats:
        o = model_util.philox4x32_10(*[np.array([c], dtype=np.uint32) for c in ctr], key[0], key[1])
        assert tuple(int(x[0]) for x in o) == exp


def test_numpy_stream_equivalences_behind_mask_to_indices():
    # choice(replace=False) == permutation[:k]; choice(replace=True) == randint
    a = np.random.RandomState(5).choice(700, 512, replace=False)
    b = np.random.RandomState(5).permutation(700)[:512]
    assert (a == b).all()
    a = np.random.RandomState(5).choice(100, 412, replace=True)
    b = np.random.RandomState(5).randint(0, 100, 412)
    assert (a == b).all()


@pytest.mark.parametrize('mode', ['numpy_legacy', 'philox'])
def test_mask_to_indices_properties(mode):
    rng = np.random.RandomState(0)
    N, M = 2048, 512
    counts = [0, 1, 5, 511, 512, 513, 2048]
    mask = np.zeros((len(counts), N), np.float32)
    for i, c in enumerate(counts):
        mask[i, rng.permutation(N)[:c]] = 1
    kw = dict(rng_mode=mode, rng=np.random.RandomState(3), seed=77)
    ind = model_util.mask_to_indices(mask, M, **kw)
    assert ind.dtype == np.int32 and ind.shape == (len(counts), M, 2)
    for i, c in enumerate(counts):
        assert (ind[i, :, 0] == i).all()
        pts = ind[i, :, 1]
        if c == 0:
            assert (pts == 0).all()                         # empty mask -> point 0 repeated
            continue
        assert mask[i, pts].all()                           # only masked-in points
        if c > M:
            assert len(set(pts.tolist())) == M              # without replacement
        else:
            assert set(pts.tolist()) == set(np.where(mask[i] > 0.5)[0].tolist())   # every point at least once
    # deterministic
    kw2 = dict(rng_mode=mode, rng=np.random.RandomState(3), seed=77)
    assert (model_util.mask_to_indices(mask, M, **kw2) == ind).all()


def test_box_pc_representation_closed_form():
    # literal restatement (surface points + normals) vs the closed form the kernels use
    g = torch.Generator().manual_seed(0)
    B, N = 3, 50
    pc = torch.randn(B, N, 6, generator=g, dtype=torch.float64)
    center = torch.randn(B, 3, generator=g, dtype=torch.float64)
    dims = torch.rand(B, 3, generator=g, dtype=torch.float64) + 0.5
    th = torch.rand(B, generator=g, dtype=torch.float64) * 6 - 3
    rep = tf_util.tf_get_box_pc_representation((center, dims, th), pc)
    d = pc[:, :, :3] - center[:, None]
    c, s = torch.cos(th)[:, None], torch.sin(th)[:, None]
    xr, zr = c * d[..., 0] - s * d[..., 2], s * d[..., 0] + c * d[..., 2]
    l, w, h = dims[:, 0:1], dims[:, 1:2], dims[:, 2:3]
    closed = torch.stack([l / 2 - xr, l / 2 + xr, h / 2 - d[..., 1], h / 2 + d[..., 1], w / 2 - zr, w / 2 + zr], dim=2)
    assert torch.allclose(rep[:, :, 6:], closed, atol=1e-12)
    assert torch.equal(rep[:, :, :6], pc)


def test_reprojection_corner_set_matches_get_3d_box():
    # tf_create_3D_box_by_vertices_multi == get_3d_box corners as a set
    g = torch.Generator().manual_seed(1)
    B = 5
    center = torch.randn(B, 3, generator=g, dtype=torch.float64)
    dims = torch.rand(B, 3, generator=g, dtype=torch.float64) + 0.5
    th = torch.rand(B, generator=g, dtype=torch.float64) * 6 - 3
    _, c1 = tf_util.tf_create_3D_box_by_vertices_multi((center, dims, th), apply_translation=True)
    c2 = model_util.get_box3d_corners_helper(center, th, dims)
    for b in range(B):
        a = sorted(map(tuple, np.round(c1[b].numpy(), 9).tolist()))
        bb = sorted(map(tuple, np.round(c2[b].numpy(), 9).tolist()))
        assert np.allclose(a, bb, atol=1e-8)


def test_synth_projection_matches_oracle_projection():
    b = synth.make_batch(4, 64, 6, seed=5)
    pts = torch.randn(4, 8, 3, dtype=torch.float64) + torch.tensor([0., 0., 4.], dtype=torch.float64)
    uv1 = synth.project_upright_camera_to_image(pts.numpy(), b['Rtilt'].astype(np.float64), b['K'].astype(np.float64))
    depth = tf_util.project_upright_camera_to_upright_depth(pts)
    uv2, _ = tf_util.project_upright_depth_to_image(depth, torch.as_tensor(b['Rtilt'], dtype=torch.float64),
                                                    torch.as_tensor(b['K'], dtype=torch.float64))
    assert np.allclose(uv1, uv2.numpy(), atol=1e-9)


def test_tiny_hand_computed_mask_and_centroid():
    pc = torch.tensor([[[1., 2., 3., 9., 9., 9.], [3., 4., 5., 9., 9., 9.], [10., 10., 10., 0., 0., 0.], [0., 0., 0., 0., 0., 0.]],
                       [[1., 1., 1., 0., 0., 0.]] * 4])
    logits = torch.tensor([[[0., 1.], [0.5, 0.6], [1., 0.], [2., 2.]],      # tie -> 0
                           [[1., 0.]] * 4])                                   # empty mask
    mask, mean, xyz, xyz1 = semisup_models.subtract_points_mean(pc, logits)
    assert mask[..., 0].tolist() == [[1., 1., 0., 0.], [0., 0., 0., 0.]]
    assert torch.allclose(mean[0, 0], torch.tensor([2., 3., 4.]))
    assert torch.allclose(mean[1, 0], torch.zeros(3))                         # divides by max(count,1)
    assert torch.allclose(xyz1[0, 2], torch.tensor([8., 7., 6.]))


def test_anchor_to_reg_first_max_and_clamp():
    center = torch.zeros(2, 3)
    dims_cls = torch.tensor([[1., 3., 3.] + [0.] * 7, [0.] * 10])            # tie -> first
    dims_reg = torch.zeros(2, 10, 3)
    dims_reg[0, 1] = torch.tensor([-10., 0.1, 0.2])                          # clamp at 1e-5
    orient_cls = torch.zeros(2, 12)
    orient_cls[0, 5] = 1.
    orient_reg = torch.arange(24, dtype=torch.float32).reshape(2, 12) * 0.01
    da = torch.as_tensor(MEAN_DIMS_ARR, dtype=torch.float32)
    oa = torch.as_tensor(np.arange(0, 2 * np.pi, 2 * np.pi / 12), dtype=torch.float32)
    c, d, o = tf_util.tf_convert_box_params_from_anchor_to_reg_format_multi((center, dims_cls, dims_reg, orient_cls, orient_reg), None, da, oa)
    assert abs(d[0, 0].item() - 1e-5) < 1e-12 and abs(d[0, 1].item() - (MEAN_DIMS_ARR[1, 1] + 0.1)) < 1e-6
    assert torch.allclose(d[1], da[0])
    assert abs(o[0].item() - (5 * 2 * np.pi / 12 + 0.05)) < 1e-6 and abs(o[1].item() - 0.12) < 1e-6


def test_intraclass_variance_empty_group_is_zero():
    dims = torch.tensor([[1., 1., 1.], [3., 1., 1.], [5., 5., 5.]])
    cls = torch.tensor([0, 0, 2])
    train = [True, True, True] + [False] * 7
    l = weak_losses.get_intraclass_variance_loss_v1(dims, cls, train, 10, True, 0.2, 'huber')
    # class 0: mean (2,1,1): errors 1,0,0,1,0,0 -> huber .5 each -> sum 1 / 6 elements; class 1 empty -> 0; class 2 -> 0
    assert abs(l.item() - (1.0 / 6.0) / 3.0) < 1e-7


def test_bn_fold_equals_unfolded_layer_and_folded_conv6():
    v = weights.make_weights_model_F(seed=3)
    b = synth.make_batch(2, 64, 6, seed=9)
    vs = VarStore(v, dtype=torch.float64)
    pc = torch.as_tensor(b['pc'], dtype=torch.float64)
    ep = {}
    with torch.no_grad(), vs.variable_scope('class_agnostic'):
        lit = semisup_models.v1_inst_seg(pc, None, None, ep, False, vs, scope='inst_seg')
        vs.literal = False
        fold = semisup_models.v1_inst_seg(pc, None, None, ep, False, vs, scope='inst_seg')
    assert torch.allclose(lit, fold, atol=1e-10)
    w, bb = weights.fold_bn(v, 'class_agnostic/inst_seg/conv1')
    y = np.maximum(b['pc'].reshape(-1, 6).astype(np.float64) @ w.astype(np.float64) + bb, 0)
    from oracle.tf_layers import conv2d
    with torch.no_grad(), vs.variable_scope('class_agnostic/inst_seg'):
        y2 = conv2d(pc, 64, [1, 6], vs, 'conv1', True, False)
    assert np.allclose(y, y2.reshape(-1, 64).numpy(), atol=1e-5)


def test_masked_max_equals_max_over_compacted_points():
    # the identity the B200 path relies on: max_n(relu(.)*mask) == max(0, max over masked-in points)
    g = torch.Generator().manual_seed(2)
    act = torch.relu(torch.randn(3, 40, 7, generator=g))
    mask = (torch.rand(3, 40, 1, generator=g) > 0.5).float()
    mask[2] = 0
    ref = (act * mask).max(dim=1).values
    for b in range(3):
        sel = act[b][mask[b, :, 0] > 0.5]
        got = sel.max(dim=0).values if sel.shape[0] else torch.zeros(7)
        assert torch.equal(ref[b], got)


def test_training_mode_bn_updates_moving_stats():
    v = weights.make_weights_model_F(seed=3)
    vs = VarStore(v)
    x = torch.randn(16, 512)
    before = vs.vars['class_agnostic/box_est/fc1/bn/moving_mean'].clone()
    from oracle.tf_layers import fully_connected
    with vs.variable_scope('class_agnostic/box_est'):
        y = fully_connected(x, 512, vs, 'fc1', True, True, bn_decay=0.5)
    after = vs.vars['class_agnostic/box_est/fc1/bn/moving_mean']
    assert not torch.allclose(before, after)
    pre = x @ vs.vars['class_agnostic/box_est/fc1/weights'] + vs.vars['class_agnostic/box_est/fc1/biases']
    assert torch.allclose(after, 0.5 * before + 0.5 * pre.mean(0), atol=1e-5)


def test_golden_model_F_fixture():
    """Committed fixture generated by tests/golden/make_golden.py from the oracle (the reference cannot
    run here, so this pins the oracle against regressions, not against TF)."""
    path = os.path.join(GOLDEN, 'model_F_tiny.npz')
    z = np.load(path)
    variables = weights.make_weights_model_F(seed=int(z['weight_seed']), with_boxpc=True)
    b = synth.make_batch(int(z['B']), int(z['N']), 6, seed=int(z['data_seed']))
    from oracle import test_semisup
    vs = VarStore(variables)
    with torch.no_grad():
        logits, ep = test_semisup.run_graph(vs, config.cfg(), torch.as_tensor(b['pc']), torch.as_tensor(b['one_hot']))
    assert np.allclose(logits.numpy(), z['logits'], atol=2e-5)
    for k in ('F2_center', 'F_size_residuals', 'F_heading_scores', 'boxpc_fit_prob', 'stage1_center'):
        assert np.allclose(ep[k].numpy(), z[k], atol=2e-5), k


# ---------------------------------------------------------------------------------------------- box_util (3D IoU) oracle
def _obox(c, s, h):
    from oracle import box_util as bu
    return bu.get_3d_box(np.array(s, dtype=np.float64), h, np.array(c, dtype=np.float64))


def test_box3d_iou_closed_forms():
    """oracle/box_util.py restates the un-shipped box_util of the reference's upstream; pinned by closed forms."""
    from oracle import box_util as bu
    a = _obox([0, 0, 0], [2, 1, 1], 0.3)
    assert np.allclose(bu.box3d_iou(a, a), (1.0, 1.0))
    # shifted by half the length along x: intersection 1 of volume 2 + 2 - 1
    assert np.allclose(bu.box3d_iou(_obox([0, 0, 0], [2, 1, 1], 0.0), _obox([1, 0, 0], [2, 1, 1], 0.0)), (1 / 3, 1 / 3))
    # the 2 x 1 footprint rotated by 90 degrees: BEV intersection 1 x 1
    assert np.allclose(bu.box3d_iou(_obox([0, 0, 0], [2, 1, 1], 0.0), _obox([0, 0, 0], [2, 1, 1], np.pi / 2)), (1 / 3, 1 / 3))
    assert np.allclose(bu.box3d_iou(_obox([0, 0, 0], [2, 1, 1], 0.0), _obox([5, 0, 0], [2, 1, 1], 0.0)), (0.0, 0.0))
    # half-height shift along y leaves the BEV IoU at 1
    assert np.allclose(bu.box3d_iou(_obox([0, 0, 0], [2, 1, 1], 0.0), _obox([0, 0.5, 0], [2, 1, 1], 0.0)), (1 / 3, 1.0))
    # containment: small box inside a big one -> vol_small / vol_big
    assert np.allclose(bu.box3d_iou(_obox([0, 0, 0], [4, 4, 4], 0.7), _obox([0.2, 0.1, -0.3], [1, 1, 1], 0.7))[0], 1 / 64)


def test_box3d_iou_symmetry_and_monte_carlo():
    from oracle import box_util as bu
    rng = np.random.RandomState(0)
    for _ in range(4):
        c1, s1, h1 = rng.uniform(-0.3, 0.3, 3), rng.uniform(0.5, 2, 3), rng.uniform(-3, 3)
        c2, s2, h2 = rng.uniform(-0.3, 0.3, 3), rng.uniform(0.5, 2, 3), rng.uniform(-3, 3)
        i3, i2 = bu.box3d_iou(_obox(c1, s1, h1), _obox(c2, s2, h2))
        j3, j2 = bu.box3d_iou(_obox(c2, s2, h2), _obox(c1, s1, h1))
        assert abs(i3 - j3) < 1e-12 and abs(i2 - j2) < 1e-12
        P = rng.uniform(-2.2, 2.2, (200000, 3))

        def inside(c, s, h):
            d = P - c
            co, si = np.cos(h), np.sin(h)
            x, z = co * d[:, 0] - si * d[:, 2], si * d[:, 0] + co * d[:, 2]           # inverse of roty
            return (np.abs(x) <= s[0] / 2) & (np.abs(d[:, 1]) <= s[2] / 2) & (np.abs(z) <= s[1] / 2)
        a, b = inside(c1, s1, h1), inside(c2, s2, h2)
        assert abs(i3 - (a & b).sum() / max((a | b).sum(), 1)) < 0.02


def test_perturb_box_to_diff_ious_bands_and_streams():
    """box_pc_fit_dataset.py:211-244: accepted IoUs lie strictly inside the band; the philox stream is reproducible and
    the numpy_legacy mode consumes numpy's global-style stream (7 uniforms per attempt)."""
    from oracle import box_pc_fit_dataset as od
    c, s, h = np.array([0.1, 0.2, 3.0]), np.array([1.9, 1.2, 0.9]), 0.4
    for band in ([0.7, 1.0], [0.01, 0.25]):
        r = od.perturb_box_to_diff_ious(c, s, h, band, rng_mode='philox', seed=9, box_index=3)
        assert band[0] < r[3] < band[1] and r[7] >= 1
        r2 = od.perturb_box_to_diff_ious(c, s, h, band, rng_mode='philox', seed=9, box_index=3)
        assert np.array_equal(r[0], r2[0]) and r[7] == r2[7]
        rng = np.random.RandomState(5)
        r3 = od.perturb_box_to_diff_ious(c, s, h, band, rng_mode='numpy_legacy', rng=rng)
        assert band[0] < r3[3] < band[1]
        ref = np.random.RandomState(5)
        ref.uniform(size=7 * r3[7])
        assert rng.uniform() == ref.uniform()                     # exactly 7 draws per attempt were consumed
        assert np.allclose(r3[0], c + r3[4]) and np.allclose(r3[1], s + r3[5]) and np.isclose(r3[2], h + r3[6])


# ----------------------------------------------------------------------------- inference post-processing
def test_oracle_label_conversions_hand_cases():
    from oracle import roi_seg_box3d_dataset as o
    from transferable3d_b200.constants import type_mean_size
    # class2angle: bin 9 of 12 = 270 deg -> wrapped to -90 deg; bin 3 + 0.1 stays
    assert abs(o.class2angle(9, 0.0, 12) - (-np.pi / 2)) < 1e-12
    assert abs(o.class2angle(3, 0.1, 12) - (np.pi / 2 + 0.1)) < 1e-12
    assert abs(o.class2angle(9, 0.0, 12, to_label_format=False) - 1.5 * np.pi) < 1e-12
    # rotate_pc_along_y by +90 deg: (x, z) = (1, 0) -> (0, 1)  [x' = c x - s z, z' = s x + c z]
    p = o.rotate_pc_along_y(np.array([[1.0, 5.0, 0.0]]), np.pi / 2)
    assert np.allclose(p, [[0.0, 5.0, 1.0]], atol=1e-12)
    # from_prediction_to_label_format: chair (class 3), zero residuals, rot_angle = pi/2, centre (0, 1, 2):
    # centre is rotated by -pi/2: (x, z) = (0, 2) -> (2, 0); ty = 1 + h/2; ry = bin 0 + rot
    h, w, l, tx, ty, tz, ry = o.from_prediction_to_label_format(np.array([0.0, 1.0, 2.0]), 0, 0.0, 3, np.zeros(3), np.pi / 2)
    ml, mw, mh = type_mean_size['chair']
    assert (h, w, l) == (mh, mw, ml)
    assert np.allclose([tx, ty, tz, ry], [2.0, 1.0 + mh / 2, 0.0, np.pi / 2], atol=1e-12)


def test_oracle_inference_scores_hand_case():
    from oracle import test_semisup as o
    # 1 frustum, 3 points: logits (0,0) tie -> class 0; (0, ln 3) -> in, p1 = 3/4; (1, 0) -> out
    logits = np.array([[[0.0, 0.0], [0.0, np.log(3.0)], [1.0, 0.0]]])
    hs = np.zeros((1, 12)); hs[0, 5] = np.log(12.0)       # softmax max = 12 / (11 + 12)
    ss = np.zeros((1, 10)); ss[0, 2] = np.log(10.0)       # 10 / 19
    hr = np.arange(12.0)[None] * 0.1
    sr = np.arange(30.0).reshape(1, 10, 3)
    seg, mmp, hc, hres, sc, sres, score = o.inference_scores(logits, hs, hr, ss, sr, fit_prob=np.array([0.5]))
    assert seg.tolist() == [[0, 1, 0]]
    assert abs(mmp[0] - 0.75 / 2.0) < 1e-12                 # / (count + 1)
    assert hc[0] == 5 and abs(hres[0] - 0.5) < 1e-12 and sc[0] == 2 and sres[0].tolist() == [6.0, 7.0, 8.0]
    want = np.log(0.375 + 0.01) + np.log(12.0 / 23.0 + 0.01) + np.log(10.0 / 19.0 + 0.01) + np.log(0.51)
    assert abs(score[0] - want) < 1e-12


# ----------------------------------------------------------------------------- detection evaluation
def test_oracle_voc_ap_and_matching_hand_cases():
    from oracle import eval_det as o, box_util as ob
    # voc_ap: PR points (rec, prec) = (0.5, 1.0), (0.5, 0.5), (1.0, 2/3) -> envelope 1.0 on [0, .5], 2/3 on (.5, 1]
    rec, prec = np.array([0.5, 0.5, 1.0]), np.array([1.0, 0.5, 2.0 / 3.0])
    assert abs(o.voc_ap(rec, prec) - (0.5 * 1.0 + 0.5 * 2.0 / 3.0)) < 1e-12
    # 11-point: recall thresholds 0..0.5 -> 1.0 (6 points), 0.6..1.0 -> 2/3 (5 points)
    assert abs(o.voc_ap(rec, prec, use_07_metric=True) - (6 * 1.0 + 5 * 2.0 / 3.0) / 11.0) < 1e-12
    # one image, two unit GT boxes 10 apart; detections: exact hit on box 0 (score .9), duplicate of box 0 (score .8,
    # false positive: already claimed), half-shifted box on box 1 (score .7, IoU 1/3 > .25: true positive), far miss (.6)
    box = lambda x: ob.get_3d_box((1.0, 1.0, 1.0), 0.0, (x, 0.0, 0.0))
    gt = {7: [box(0.0), box(10.0)]}
    pred = {7: [(box(0.0), 0.9), (box(0.05), 0.8), (box(10.5), 0.7), (box(50.0), 0.6)]}
    rec, prec, ap, (tp, fp, ov) = o.eval_det_cls(pred, gt, 0.25, return_match=True)
    assert tp.tolist() == [1, 0, 1, 0] and fp.tolist() == [0, 1, 0, 1]
    assert abs(ov[2] - 1.0 / 3.0) < 1e-9 and ov[3] == 0.0
    assert np.allclose(rec, [0.5, 0.5, 1.0, 1.0]) and np.allclose(prec, [1.0, 0.5, 2.0 / 3.0, 0.5])
    assert abs(ap - (0.5 + 0.5 * 2.0 / 3.0)) < 1e-12


# ----------------------------------------------------------------------------- surface loss
def test_oracle_surface_distance_closed_forms():
    """Axis-aligned unit box at the origin: a point on the ray through the +x face at (a, 0, 0), a > 0, is |a - 0.5| from
    that face along the ray; the uncleaned minimum also sees the opposite face (distance a + 0.5) and the side faces
    (ray parallel: |r - r * 0.5 / 1e-5|, huge) -- so d = |a - 0.5|; rotating box and point together changes nothing."""
    from oracle import tf_util as OT
    for theta in (0.0, 0.7):
        c, s = np.cos(theta), np.sin(theta)
        box = (torch.zeros(1, 3, dtype=torch.float64), torch.ones(1, 3, dtype=torch.float64), torch.tensor([theta], dtype=torch.float64))
        for a in (0.2, 0.5, 1.3):
            # box-frame point (a,0,0) -> world: R (a,0,0) = (c a, 0, -s a)
            p = torch.tensor([[[c * a, 0.0, -s * a]]], dtype=torch.float64)
            d = OT.tf_distance_to_closest_3D_box_surface_multi(p, box)
            assert abs(float(d) - abs(a - 0.5)) < 1e-4, (theta, a, float(d))
    # dims order (l, w, h): x uses l, y uses h, z uses w
    box = (torch.zeros(1, 3, dtype=torch.float64), torch.tensor([[2.0, 4.0, 6.0]], dtype=torch.float64), torch.zeros(1, dtype=torch.float64))
    pts = torch.tensor([[[1.5, 0, 0], [0, 3.5, 0], [0, 0, 2.5]]], dtype=torch.float64)
    d = OT.tf_distance_to_closest_3D_box_surface_multi(pts, box)[0]
    assert np.allclose(d.numpy(), [0.5, 0.5, 0.5], atol=1e-4)
