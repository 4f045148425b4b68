"""-m gpu: the split-precision (f16x2) fused kernels -- csrc/chain_x2.cuh, csrc/seg_stage2_x2.cuh -- through the C ABI.

Bar: the fp32-mode bar of the north star (1e-4 of the tensor scale against the fp32 oracle), because this is the mode that
has to meet the mask-exactness target; unit chains are compared with a float64 restatement at 2e-5.
"""
import numpy as np
import pytest
import torch

from util import model_F_setup, oracle_model_F, err_stats

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from transferable3d_b200 import runtime as rt, semisup_models as sm, test_semisup as ts, weights

DEV = 'cuda:0'

CHAINS = {
    # kind -> (scope, layer names, cin)
    'seg1': ('class_agnostic/inst_seg', ['conv1', 'conv2', 'conv3', 'conv4', 'conv5'], 6),
    'tnet': ('class_agnostic/tnet', ['conv-reg1-stage1', 'conv-reg2-stage1', 'conv-reg3-stage1'], 3),
    'box': ('class_agnostic/box_est', ['conv-reg1', 'conv-reg2', 'conv-reg3', 'conv-reg4'], 3),
}


def scale_close(got, ref, tol, what):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    s = err_stats(got, ref)
    assert np.isfinite(got).all(), what
    assert s['max_abs'] <= tol * max(s['ref_scale'], 1e-6), (what, s)


def chain_ref64(variables, scope, layers, x):
    """float64 restatement of conv(1x1)+BN(eval, folded)+ReLU ... max over points. x: (B, n, cin)."""
    a = x.astype(np.float64)
    for name in layers:
        w, b = weights.fold_bn(variables, scope + '/' + name)
        a = np.maximum(a @ w.astype(np.float64) + b.astype(np.float64), 0)
    return a.max(axis=1)


@pytest.fixture(scope='module')
def store():
    variables = weights.make_weights_model_F()
    st = rt.VariableStore(variables, DEV)
    return variables, st


@pytest.mark.parametrize('kind', ['seg1', 'tnet', 'box'])
@pytest.mark.parametrize('B,N', [(3, 300), (40, 2048)])
def test_x2_chain_vs_float64(kind, B, N, built_lib, store):
    """dense chains; (3, 300): ragged last tile, one tile per CTA; (40, 2048): 640 tiles on 148 SMs = several iterations per
    CTA, ring-phase wrap, running max carried across the tiles of a frustum and flushed when the frustum changes"""
    variables, st = store
    scope, layers, cin = CHAINS[kind]
    rng = np.random.RandomState(B + N)
    pc = rng.randn(B, N, 6).astype(np.float32)
    code = {'seg1': rt.CHAIN_SEG1, 'tnet': rt.CHAIN_TNET, 'box': rt.CHAIN_BOX}[kind]
    arena = st.chain_arena(scope, code, layers, x2=True)
    out = rt.chain_max(code, torch.as_tensor(pc).to(DEV), arena, x2=True)
    ref = chain_ref64(variables, scope, layers, pc[:, :, :cin])
    scale_close(out.cpu().numpy(), ref, 2e-5, kind)


def test_x2_masked_chain_center_and_compaction(built_lib, store):
    """tnet / box chains on compacted points with a subtracted centre: counts 0, 1, 127, 128, 129, 1000, 2048"""
    variables, st = store
    rng = np.random.RandomState(5)
    counts = [0, 1, 127, 128, 129, 1000, 2048]
    B, N = len(counts), 2048
    pc = rng.randn(B, N, 6).astype(np.float32)
    center = rng.randn(B, 3).astype(np.float32)
    idx = np.zeros((B, N), np.int32)
    for i, c in enumerate(counts):
        idx[i, :c] = np.sort(rng.permutation(N)[:c])
    scope, layers, _ = CHAINS['box']
    arena = st.chain_arena(scope, rt.CHAIN_BOX, layers, x2=True)
    out = rt.chain_max(rt.CHAIN_BOX, torch.as_tensor(pc).to(DEV), arena, center=torch.as_tensor(center).to(DEV),
                       idx=torch.as_tensor(idx).to(DEV), count=torch.as_tensor(np.asarray(counts, np.int32)).to(DEV), x2=True)
    out = out.cpu().numpy()
    assert np.abs(out[0]).max() == 0.0                                   # empty mask -> zero feature
    for i, c in enumerate(counts[1:], start=1):
        x = pc[i, idx[i, :c], :3] - center[i]
        ref = chain_ref64(variables, scope, layers, x[None])[0]
        scale_close(out[i], ref, 2e-5, 'box count %d' % c)


def test_x2_boxpc_chain(built_lib, store):
    from oracle import tf_util as otu
    variables, st = store
    rng = np.random.RandomState(9)
    B, N = 5, 700
    pc = rng.randn(B, N, 6).astype(np.float32)
    box = (rng.randn(B, 3).astype(np.float32), (0.5 + rng.rand(B, 3)).astype(np.float32), rng.uniform(-3, 3, B).astype(np.float32))
    scope, layers = 'D_boxpc_branch/box_pc_mask_model', ['conv-reg1', 'conv-reg2', 'conv-reg3', 'conv-reg4']
    arena = st.chain_arena(scope, rt.CHAIN_BOXPC, layers, x2=True)
    out = rt.chain_max(rt.CHAIN_BOXPC, torch.as_tensor(pc).to(DEV), arena, box=tuple(torch.as_tensor(t).to(DEV) for t in box), x2=True)
    rep = otu.tf_get_box_pc_representation(tuple(torch.as_tensor(t).double() for t in box), torch.as_tensor(pc).double()).numpy()
    ref = chain_ref64(variables, scope, layers, rep)
    scale_close(out.cpu().numpy(), ref, 5e-5, 'boxpc')      # the fp32 plane distances of the prologue are part of the error


@pytest.fixture(scope='module')
def setup():
    variables, batch, FLAGS, info = model_F_setup(4)
    ologits, oep = oracle_model_F(variables, batch, FLAGS)
    st = rt.VariableStore(variables, DEV)
    return dict(variables=variables, batch=batch, FLAGS=FLAGS, ologits=ologits, oep=oep, store=st,
                pc=torch.as_tensor(batch['pc']).to(DEV), oh=torch.as_tensor(batch['one_hot']).to(DEV))


def test_x2_model_F_vs_oracle(built_lib, setup):
    """model F + one BoxPC refine in f16x2: logits at the fp32-mode bar; everything downstream against the oracle
    continued from the GPU's own mask (identical masks by construction)."""
    rt.set_default_store(setup['store'])
    with rt.precision('f16x2'), torch.no_grad():
        logits, ep = ts.build_graph(setup['FLAGS'], setup['pc'], setup['oh'])
    scale_close(logits.cpu().numpy(), setup['ologits'].numpy(), 1e-4, 'logits')
    gmask = (logits[..., 0] < logits[..., 1]).float().cpu().numpy()
    agree = (gmask == (setup['ologits'][..., 0] < setup['ologits'][..., 1]).float().numpy()).mean()
    assert agree > 0.9995, agree
    _, oep = oracle_model_F(setup['variables'], setup['batch'], setup['FLAGS'], oracle_mask=gmask)
    for k in ('stage1_center', 'feats_lv1', 'box_params', 'F_center', 'F_heading_scores', 'F_heading_residuals', 'F_size_scores',
              'F_size_residuals', 'F2_center', 'F2_heading_residuals', 'F2_size_residuals', 'boxpc_fit_prob'):
        scale_close(ep[k].cpu().numpy(), oep[k].numpy(), 2e-4, k)


def test_x2_debias_changes_only_the_scales(built_lib, setup):
    """the truncation correction is a per-layer factor (1 + c n): switching it off moves the logits by ~1e-6 relative,
    not more (a wrong step count or a correction applied twice would show here)"""
    rt.set_default_store(setup['store'])
    old = rt.get_x2_debias()
    try:
        with rt.precision('f16x2'), torch.no_grad():
            a = sm.v1_inst_seg(setup['pc'], None, None, {}, False, scope='class_agnostic/inst_seg').cpu().numpy()
            rt.set_x2_debias(0.0)
            b = sm.v1_inst_seg(setup['pc'], None, None, {}, False, scope='class_agnostic/inst_seg').cpu().numpy()
    finally:
        rt.set_x2_debias(old)
    s = err_stats(a, b)
    assert 0 < s['max_abs'] <= 5e-5 * s['ref_scale'], s


def test_x2_properties(built_lib, setup):
    """size-independent properties: point-order invariance, batch independence, duplicated points"""
    st, pc = setup['store'], setup['pc']
    arena = st.chain_arena('class_agnostic/inst_seg', rt.CHAIN_SEG1, ['conv1', 'conv2', 'conv3', 'conv4', 'conv5'], x2=True)
    g1 = rt.chain_max(rt.CHAIN_SEG1, pc, arena, x2=True)
    perm = torch.randperm(2048, device=DEV)
    assert torch.equal(g1, rt.chain_max(rt.CHAIN_SEG1, pc[:, perm].contiguous(), arena, x2=True))
    assert torch.equal(g1[1:3], rt.chain_max(rt.CHAIN_SEG1, pc[1:3].contiguous(), arena, x2=True))
    half = rt.chain_max(rt.CHAIN_SEG1, pc[:, :1024].contiguous(), arena, x2=True)
    dup = torch.cat([pc[:, :1024], pc[:, :1024]], dim=1).contiguous()
    assert torch.equal(rt.chain_max(rt.CHAIN_SEG1, dup, arena, x2=True), half)


@pytest.mark.parametrize('mode', ['f16x2'])
def test_x2_session_cuda_graph_replay_equals_eager(built_lib, setup, mode):
    b, FLAGS = setup['batch'], setup['FLAGS']
    fetch = ['logits', 'F2_center', 'F2_heading_residuals', 'F2_size_residuals', 'boxpc_fit_prob']
    with rt.precision(mode):
        sess_g, ops = ts.get_model(4, 2048, 6, FLAGS=FLAGS, variables=setup['store'], cuda_graph=True)
        sess_e, _ = ts.get_model(4, 2048, 6, FLAGS=FLAGS, variables=setup['store'], cuda_graph=False)
        for rep in range(2):
            feed = {ops['pc_pl']: np.roll(b['pc'], rep, axis=0), ops['one_hot_vec_pl']: np.roll(b['one_hot'], rep, axis=0),
                    ops['is_training_pl']: False}
            got = sess_g.run(fetch, feed)
            ref = sess_e.run(fetch, feed)
            torch.cuda.synchronize()
            for k, g, r in zip(fetch, got, ref):
                assert torch.equal(g, r), (mode, rep, k)
