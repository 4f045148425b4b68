"""-m gpu: the BoxPC variants of SURVEY 8(f) rank 3 against the oracle:

  * representation B (semisup_models.independent_box_pc_mask_features_model, semisup_models.py:400-471) -- eval graph in all
    precision modes, and the train_boxpc step (loss, gradients, moving statistics) with BOX_PC_MASK_REPRESENTATION = 'B';
  * the NORMALIZE_PC options (normalize_pc = True, 'SD' / 'Spread', semisup_models.py:335-343, 413-421) and the mask input
    of both representations; t3d_normalize_pc against models/tf_util.py:134-173;
  * class-confidence weighting of the BoxPC deltas / delta loss (boxpc_sunrgbd.py:76-91, 166-177): every flag combination,
    loss and gradients of the training step.
"""
import numpy as np
import pytest
import torch

from util import err_stats, boxpc_seed_without_pool_ties, assert_grad_close

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from transferable3d_b200 import runtime as rt, semisup_models as sm, boxpc_sunrgbd as bp, tf_util as tu, train_boxpc as tb
    from transferable3d_b200 import weights, synth, config

DEV = 'cuda:0'


def _box_and_pc(B, N, seed):
    rng = np.random.RandomState(seed)
    pc = np.concatenate([rng.randn(B, N, 3) * [1.0, 0.5, 1.5] + [0.3, -0.2, 4.0], rng.rand(B, N, 3)], axis=2).astype(np.float32)
    box = (pc[:, :, :3].mean(1) + rng.randn(B, 3).astype(np.float32) * 0.2,
           (rng.rand(B, 3) * 1.5 + 0.5).astype(np.float32), (rng.rand(B) * 6.28).astype(np.float32))
    one_hot = np.eye(10, dtype=np.float32)[rng.randint(0, 10, B)]
    mask = (rng.rand(B, N, 1) < 0.5).astype(np.float32)
    return pc, box, one_hot, mask


@pytest.mark.parametrize('mode', [0, 1])
def test_normalize_pc_kernel(mode, built_lib):
    from oracle import tf_util as otu
    pc, _, _, _ = _box_and_pc(5, 777, 3)
    fn = (tu.tf_normalize_point_clouds_to_mean_zero_and_unit_var, tu.tf_normalize_point_clouds_to_01)[mode]
    ofn = (otu.tf_normalize_point_clouds_to_mean_zero_and_unit_var, otu.tf_normalize_point_clouds_to_01)[mode]
    for x in (pc, pc[:, :, :3].copy()):                      # 6 channels (rgb copied) and xyz only
        got = fn(torch.as_tensor(x).to(DEV)).cpu().numpy()
        ref = ofn(torch.as_tensor(x).double()).numpy()
        assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
        if x.shape[2] == 6:
            assert np.array_equal(got[:, :, 3:], x[:, :, 3:])


def _run_model(rep, variables, pc, box, one_hot, mask, normalize, method, precision):
    from oracle.tf_layers import VarStore
    from oracle import semisup_models as osm
    FLAGS = config.cfg(BOX_PC_MASK_REPRESENTATION=rep)
    vs = VarStore(variables)
    T = torch.as_tensor
    with torch.no_grad():
        oout, ofeats = osm.box_pc_mask_features_model(tuple(T(b) for b in box), T(pc), None if mask is None else T(mask), 9, False, {},
                                                      False, False, vs, normalize_pc=normalize, normalize_method=method,
                                                      one_hot_vec=T(one_hot), c=FLAGS, scope='box_pc_mask_model')
    store = rt.VariableStore(variables, DEV)
    rt.set_default_store(store)
    with rt.precision(precision), torch.no_grad():
        out, feats = sm.box_pc_mask_features_model(tuple(T(b).to(DEV) for b in box), T(pc).to(DEV),
                                                   None if mask is None else T(mask).to(DEV), 9, False, {}, False, False,
                                                   normalize_pc=normalize, normalize_method=method, one_hot_vec=T(one_hot).to(DEV),
                                                   c=FLAGS, scope='box_pc_mask_model')
    torch.cuda.synchronize()
    return out.cpu().numpy(), {k: v.cpu().numpy() for k, v in feats.items()}, oout.numpy(), {k: v.numpy() for k, v in ofeats.items()}


@pytest.mark.parametrize('rep', ['A', 'B'])
@pytest.mark.parametrize('normalize,method,masked', [(False, 'SD', False), (True, 'SD', False), (True, 'Spread', True), (False, 'SD', True)])
def test_boxpc_model_variants_fp32(rep, normalize, method, masked, built_lib):
    B, N = 6, 512
    pc, box, one_hot, mask = _box_and_pc(B, N, 11)
    v = weights.make_weights_boxpc(use_one_hot=True, rep=rep)
    if masked:      # an extra mask channel widens conv-reg1 of representation A (semisup_models.py:345-350)
        if rep == 'A':
            rng = np.random.RandomState(5)
            w = v['box_pc_mask_model/conv-reg1/weights']
            v['box_pc_mask_model/conv-reg1/weights'] = np.concatenate([w, rng.randn(1, 1, 1, 128).astype(np.float32) * 0.1], axis=1)
    out, feats, oout, ofeats = _run_model(rep, v, pc, box, one_hot, mask if masked else None, normalize, method, 'fp32')
    for k in ofeats:
        assert feats[k].shape == ofeats[k].shape
        assert np.abs(feats[k] - ofeats[k]).max() <= 1e-4 * max(1.0, np.abs(ofeats[k]).max()), k
    assert np.abs(out - oout).max() <= 1e-4 * max(1.0, np.abs(oout).max())


@pytest.mark.parametrize('precision,tol', [('f16x2', 1e-4), ('bf16', 3e-2)])
def test_boxpc_rep_b_fused_chain(precision, tol, built_lib):
    """Representation B through the fused chain kernel (CHAIN_BOXPCB): B = 300 frustums so that CTAs run several tiles."""
    B, N = 300, 1024
    pc, box, one_hot, _ = _box_and_pc(B, N, 21)
    v = weights.make_weights_boxpc(use_one_hot=True, rep='B')
    out, feats, oout, ofeats = _run_model('B', v, pc, box, one_hot, None, False, 'SD', precision)
    k = 'box_pc_mask_model_feats_lv1'
    assert feats[k].shape == (B, 1024)
    s = err_stats(feats[k], ofeats[k])
    assert s['max_abs'] <= tol * max(1.0, np.abs(ofeats[k]).max()), s
    assert np.abs(out - oout).max() <= 3 * tol * max(1.0, np.abs(oout).max())


def _train_setup(rep, B, N, seed=3, **flags):
    v = weights.make_weights_boxpc(rep=rep)
    feed = synth.make_boxpc_batch(B, N, 6, seed=seed)
    rng = np.random.RandomState(seed)
    names = ('dp1', 'dp2') if rep == 'A' else ('dp2', 'dp3')
    masks = {names[0]: (rng.rand(B, 512) < 0.7).astype(np.float32), names[1]: (rng.rand(B, 256) < 0.7).astype(np.float32)}
    return v, feed, masks, config.cfg(BOXPC_WEIGHT_DELTA=4., BOX_PC_MASK_REPRESENTATION=rep, **flags)


def _check_step(v, feed, masks, FLAGS, B, N):
    from oracle import train_boxpc as otb
    oloss, ograds, ovs, oep = otb.loss_and_grads(v, FLAGS, feed, masks, global_step=0)
    _, ograds64, _, _ = otb.loss_and_grads(v, FLAGS, feed, masks, global_step=0, dtype=torch.float64)
    g = tb.BoxPCTrainGraph(v, FLAGS, B, N, 6, DEV)
    out = g.forward_backward(feed, masks)
    torch.cuda.synchronize()
    assert abs(float(out['loss']) - float(oloss)) <= 1e-4 * max(1.0, abs(float(oloss)))
    assert np.abs(out['boxpc_delta_center'].cpu().numpy() - oep['boxpc_delta_center'].detach().numpy()).max() <= 1e-4
    wscale = {}
    for name, og in ograds.items():
        got = g.grad[name[len('box_pc_mask_model/'):]].cpu().numpy().reshape(-1)
        ref, ref64 = og.numpy().reshape(-1), ograds64[name].numpy().reshape(-1)
        layer = name.rsplit('/', 1)[0] if not name.endswith(('gamma', 'beta')) else name.rsplit('/', 2)[0]
        if name.endswith('weights'):
            wscale[layer] = float(np.abs(ref).mean())
        assert_grad_close(name, got, ref, ref64, scale_floor=1e-2 * wscale.get(layer, 0.0))
    for k, mv in g.moving.items():
        ref = ovs.vars['box_pc_mask_model/' + k].numpy()
        small_batch = k.endswith('variance') and ('fc' in k)
        s = err_stats(mv.cpu().numpy(), ref)
        assert s['max_abs'] <= (2e-2 if small_batch else 2e-4) * max(s['ref_scale'], 1e-3), (k, s)
    return g


@pytest.mark.parametrize('B,N', [(8, 256), (16, 1024)])
def test_boxpc_rep_b_train_step_vs_oracle(B, N, built_lib):
    seed = boxpc_seed_without_pool_ties(lambda s: _train_setup('B', B, N, seed=s), B, N, conv_last=7)
    v, feed, masks, FLAGS = _train_setup('B', B, N, seed=seed)
    g = _check_step(v, feed, masks, FLAGS, B, N)
    assert 'extract_box_feats/fc0/weights' in g.grad and float(g.grad['extract_box_feats/fc0/weights'].abs().sum()) > 0
    losses = [float(g.step(feed, masks)['loss']) for _ in range(10)]
    assert np.isfinite(losses).all() and losses[-1] < 0.8 * losses[0], losses


@pytest.mark.parametrize('flags', [dict(BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF=True, BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA=True),
                                   dict(BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF=True, BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA=False),
                                   dict(BOXPC_WEIGH_DELTA_LOSS_BY_CLS_GT=True),
                                   dict(BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF=True, BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA=True),
                                   dict(BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF=True, BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA=False),
                                   dict(BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF=True, BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF=True,
                                        BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA=False, BOXPC_DELTA_LOSS_TYPE='mse')])
def test_boxpc_weighted_delta_loss_train_step(flags, built_lib):
    seed = boxpc_seed_without_pool_ties(lambda s: _train_setup('A', 8, 256, seed=s, **flags), 8, 256, conv_last=3)
    v, feed, masks, FLAGS = _train_setup('A', 8, 256, seed=seed, **flags)
    _check_step(v, feed, masks, FLAGS, 8, 256)


def test_boxpc_weighted_delta_loss_host_mirror(built_lib):
    """boxpc_sunrgbd.get_loss / get_boxpc_delta_loss with the weighting flags on given predictions."""
    from oracle import boxpc_sunrgbd as obp
    rng = np.random.RandomState(0)
    B = 37
    logits = rng.randn(B, 2).astype(np.float32)
    deltas = (rng.randn(B, 3).astype(np.float32), rng.randn(B, 3).astype(np.float32) * 2, rng.randn(B).astype(np.float32))
    labels = (rng.rand(B).astype(np.float32), (rng.randn(B, 3).astype(np.float32), rng.randn(B, 3).astype(np.float32),
                                               rng.randn(B).astype(np.float32)))
    T = torch.as_tensor
    rt.set_default_store(rt.VariableStore({}, DEV))
    for flags in (dict(BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF=True), dict(BOXPC_WEIGH_DELTA_LOSS_BY_CLS_GT=True), dict()):
        FLAGS = config.cfg(**flags)
        ol = T(logits)
        oep = {'logits_for_weigh': torch.softmax(ol, dim=1)[:, 1]}
        opred = (ol, tuple(T(d) for d in deltas))
        olab = (T(labels[0]), tuple(T(d) for d in labels[1]))
        ref_total = obp.get_loss(opred, olab, oep, c=FLAGS)
        ref_delta = obp.get_boxpc_delta_loss(opred, olab, oep, reduce_loss=False, c=FLAGS)
        pred = (T(logits).to(DEV), tuple(T(d).to(DEV) for d in deltas))
        lab = (T(labels[0]).to(DEV), tuple(T(d).to(DEV) for d in labels[1]))
        got_total = bp.get_loss(pred, lab, {}, c=FLAGS)
        got_delta = bp.get_boxpc_delta_loss(pred, lab, {}, reduce_loss=False, c=FLAGS)
        assert abs(float(got_total) - float(ref_total)) <= 1e-5 * max(1.0, abs(float(ref_total)))
        assert np.abs(got_delta.cpu().numpy() - ref_delta.numpy()).max() <= 1e-5
