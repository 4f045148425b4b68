"""The cases that pin `oracle/` against the reference's own source.  Each case runs the SAME seeded inputs, weights, flags and
dropout masks through (a) the reference's files executed on tests/golden/tf1_shim.py (`reference`, needs /root/reference) and
(b) the oracle (`oracle`), both in float64, and returns a flat {name: array} dict of what is compared.  Large tensors
(gradients, point-wise features) are compared through `sig`: norm, sum and four seeded random projections.

make_reference_golden.py stores (a) under tests/golden/ref_<case>.npz; tests/test_oracle_vs_reference_cpu.py compares (b) with
the stored file everywhere, and re-runs (a) against the file where the reference tree is present.
TEST INFRASTRUCTURE: nothing under transferable3d_b200/ imports this.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import reference_runner as rr  # noqa: E402
from transferable3d_b200 import weights, synth, config  # noqa: E402

F64 = torch.float64
FULL_LIMIT = 1024          # tensors up to this many elements are stored whole


def sig(a):
    a = np.asarray(a, dtype=np.float64).ravel()
    rng = np.random.RandomState(a.size % (2 ** 31 - 1))
    proj = rng.standard_normal((4, a.size)) @ a if a.size else np.zeros(4)
    return np.concatenate([[np.linalg.norm(a), a.sum()], proj])


def pack(out, name, value):
    """Adds `value` (array / tensor / T / tuple of them) to the flat dict: whole when small, as a signature otherwise."""
    value = rr.to_np(value)
    if value is None:
        return
    if isinstance(value, (list, tuple)):
        for i, v in enumerate(value):
            pack(out, '%s.%d' % (name, i), v)
        return
    if isinstance(value, dict):
        for k in sorted(value):
            pack(out, '%s.%s' % (name, k), value[k])
        return
    a = np.asarray(value)
    if a.dtype == object or a.dtype.kind in 'US':
        return
    a = a.astype(np.float64)
    if a.size <= FULL_LIMIT:
        out[name] = a
    else:
        out[name + '#sig'] = sig(a)


def pack_end_points(out, ep, prefix='ep.'):
    for k in sorted(ep):
        if k.startswith('_') or k in ('intraclsdims_train_classes', 'inactive_vol_train_classes'):
            continue
        pack(out, prefix + k, ep[k])


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


# ---- case 1: the test graph of model F (test_semisup.get_model, test_semisup.py:61-180, called as is) -------------------
def _model_F_inputs(B=4, N=128):
    return weights.make_weights_model_F(seed=11), synth.make_batch(B, N, 6, seed=2024)


def _model_F_reference(refine, mask_pc):
    v, b = _model_F_inputs()
    B, N = b['pc'].shape[:2]
    out = {}
    with rr.Reference() as R:
        ts = R.mod('test_semisup')
        FLAGS = R.flags(use_one_hot=True, refine=refine, mask_pc_for_boxpc=mask_pc, SEMI_MODEL='F', BOX_PC_MASK_REPRESENTATION='A')
        ts.FLAGS, ts.MODEL, ts.GPU_INDEX, ts.MODEL_PATH = FLAGS, R.mod('semisup_v1_sunrgbd'), 0, None
        # placeholders in creation order: is_training, then semisup_v1_sunrgbd.placeholder_inputs (pc, bg_pc, img, one_hot, ...)
        R.reset(v, feeds=[False, b['pc'], None, None, b['one_hot']] + [None] * 14)
        R.quiet()
        _, ops = ts.get_model(B, N, 6)
        R.quiet(False)
        pack(out, 'logits', ops['logits'])
        pack_end_points(out, ops['end_points'])
    return out


def _model_F_oracle(refine, mask_pc):
    from oracle.tf_layers import VarStore
    from oracle import test_semisup as ots
    v, b = _model_F_inputs()
    FLAGS = config.cfg(refine=refine, mask_pc_for_boxpc=mask_pc)
    vs = VarStore(v, dtype=F64)
    with torch.no_grad():
        logits, ep = ots.run_graph(vs, FLAGS, torch.as_tensor(b['pc']).to(F64), torch.as_tensor(b['one_hot']).to(F64))
    out = {}
    pack(out, 'logits', logits)
    pack_end_points(out, ep)
    return out


# ---- case 2: the BoxPC training graph (train_boxpc.train(), graph block executed from the script's AST) -----------------
def _boxpc_inputs(rep, B=6, N=96, seed=3):
    v = weights.make_weights_boxpc(rep=rep)
    feed = synth.make_boxpc_batch(B, N, 6, seed=seed)
    rng = np.random.RandomState(seed)
    names = ('dp1', 'dp2') if rep == 'A' else ('dp2', 'dp3')
    masks = {names[0]: (rng.rand(B, 512) < 0.7).astype(np.float32), names[1]: (rng.rand(B, 256) < 0.7).astype(np.float32)}
    return v, feed, masks


LEAN = [False]      # flag-variant cases keep the loss, the trained-variable list and the gradient signatures only


def _pack_step(out, loss, grads, moving, ep=None):
    if LEAN[0]:
        moving = {}
    out['loss'] = np.asarray([float(loss)])
    out['n_trained'] = np.asarray([float(len(grads))])
    for k in sorted(grads):
        g = grads[k]
        out['has_grad.' + k] = np.asarray([0.0 if g is None else 1.0])
        if g is not None:
            out['grad.' + k + '#sig'] = sig(np.asarray(g))
    for k in sorted(moving):
        pack(out, 'moving.' + k, moving[k])
    if ep is not None:
        pack_end_points(out, ep)


def _boxpc_train_reference(rep, flags):
    v, feed, masks = _boxpc_inputs(rep)
    B, N = feed['pc'].shape[:2]
    out = {}
    with rr.Reference() as R:
        tf = R.tf
        FLAGS = R.flags(use_one_hot=False, NUM_CHANNELS=6, restore_model_path=None, BOX_PC_MASK_REPRESENTATION=rep, **flags)
        # boxpc_sunrgbd.placeholder_inputs creates y_orient_delta before y_dims_delta (boxpc_sunrgbd.py:46-49)
        feeds = [feed[k] if k else None for k in ('pc', 'one_hot', None, 'x_center', 'x_orient_cls', 'x_orient_reg', 'x_dims_cls',
                                                  'x_dims_reg', 'y_box_iou', 'y_center_delta', 'y_orient_delta', 'y_dims_delta')] + [True]
        st = R.reset(v, requires_grad=True, dropout_masks={'box_pc_mask_model/' + k: m for k, m in masks.items()}, feeds=feeds)
        R.quiet()
        ns = rr.exec_train_graph(R, 'train_boxpc.py', dict(
            FLAGS=FLAGS, boxpc_sunrgbd=R.mod('boxpc_sunrgbd'), BATCH_SIZE=B, NUM_POINT=N, GPU_INDEX=0, BASE_LEARNING_RATE=0.001,
            DECAY_STEP=800000, DECAY_RATE=0.5, OPTIMIZER='adam', MOMENTUM=0.9, BN_DECAY_DECAY_STEP=800000.))
        R.quiet(False)
        op = ns['train_op']
        assert op.var_list is None and op.loss is ns['loss']
        tv = [x for x in tf.get_collection(tf.GraphKeys.TRAINABLE_VARIABLES) if x.t.is_floating_point()]
        grads = torch.autograd.grad(ns['loss'].t, [x.t for x in tv], allow_unused=True)
        moving = {k: _np(var.t) for k, var in st.vars.items() if 'moving' in k}
        _pack_step(out, ns['loss'].t.detach(), {x.name[:-2]: _np(g) for x, g in zip(tv, grads)}, moving)
        out['learning_rate'] = np.asarray([float(op.optimizer.learning_rate.t)])
        out['bn_decay'] = np.asarray([float(ns['bn_decay'].t)])
        for k in ('boxpc_fit_logits', 'boxpc_delta_center', 'boxpc_delta_size', 'boxpc_delta_angle', 'logits_for_weigh', 'pred_boxpc_fit'):
            pack(out, 'ep.' + k, ns['end_points'][k])
    return out


def _boxpc_train_oracle(rep, flags):
    from oracle import train_boxpc as otb
    v, feed, masks = _boxpc_inputs(rep)
    B = feed['pc'].shape[0]
    FLAGS = config.cfg(BOX_PC_MASK_REPRESENTATION=rep, **flags)
    loss, grads, vs, ep = otb.loss_and_grads(v, FLAGS, feed, masks, global_step=0, dtype=F64)
    out = {}
    moving = {k: _np(t) for k, t in vs.vars.items() if 'moving' in k}
    _pack_step(out, loss, {k: _np(g) for k, g in grads.items()}, moving)
    out['learning_rate'] = np.asarray([otb.get_learning_rate(0, B)])
    out['bn_decay'] = np.asarray([otb.get_bn_decay(0, B)])
    for k in ('boxpc_fit_logits', 'boxpc_delta_center', 'boxpc_delta_size', 'boxpc_delta_angle', 'logits_for_weigh', 'pred_boxpc_fit'):
        pack(out, 'ep.' + k, ep[k])
    return out


# ---- cases 3 / 4: the semi-supervised training graphs (train_semisup_adv.train() / train_semisup.train()) ------------------
CFG5 = dict(SEMI_TRAIN_BOX_TRAIN_CLASS_AG_TNET=True, SEMI_TRAIN_BOX_TRAIN_CLASS_AG_BOX=True, SEMI_BOXPC_MIN_FIT_LOSS_AFT_REFINE=True,
            WEAK_WEIGHT_INTRACLASSVAR=2., WEAK_WEIGHT_REPROJECTION=0.01, WEAK_REPROJECTION_ONLY_ON_2D_CLS=True,
            SEMI_MULTIPLIER_FOR_WEAK_LOSS=0.05, SEMI_WEIGHT_BOXPC_FIT_LOSS=1.)
CFG_A = dict(WEAK_WEIGHT_REPROJECTION=0.01, SEMI_MULTIPLIER_FOR_WEAK_LOSS=0.05)
SEMI_FEED_ORDER = ['pc', None, None, 'one_hot', 'labels', 'centers', 'y_orient_cls', 'y_orient_reg', 'y_dims_cls', 'y_dims_reg', None, None,
                   'Rtilt', 'K', 'rot_frust', 'box2D', 'img_dim', 'is_data_2D']       # semisup_v1_sunrgbd.py:37-64, creation order


def _semi_inputs(model, B=8, N=64, seed=11):
    v = weights.make_weights_model_F() if model == 'F' else weights.make_weights_model_A()
    feed = synth.make_batch(B, N, 6, seed=seed, is_data_2D=(np.arange(B) % 2))
    rng = np.random.RandomState(seed)
    if model == 'F':
        masks = {'class_agnostic/inst_seg/dp1': (rng.rand(B, N, 128) < 0.5).astype(np.float32),
                 'class_dependent/box_refine/dp0': (rng.rand(B, 512) < 0.5).astype(np.float32),
                 'class_dependent/box_refine/dp1': (rng.rand(B, 256) < 0.5).astype(np.float32)}
    else:
        masks = {'inst_seg/dp1': (rng.rand(B, N, 128) < 0.5).astype(np.float32)}
    return v, feed, masks


SEMI_EP_KEYS_F = ('stage1_center', 'F_center', 'F_heading_scores', 'F_heading_residuals', 'F_size_scores', 'F_size_residuals',
                  'boxpc_fit_prob', 'boxpc_delta_center', 'boxpc_delta_size', 'boxpc_delta_angle', 'F2_center', 'F2_heading_residuals',
                  'F2_size_residuals', 'center', 'heading_scores', 'size_residuals')
SEMI_EP_KEYS_A = ('stage1_center', 'center', 'heading_scores', 'heading_residuals', 'size_scores', 'size_residuals', 'soft_mask')


def _semi_train_reference(model, flags):
    v, feed, masks = _semi_inputs(model)
    B, N = feed['pc'].shape[:2]
    script = 'train_semisup_adv.py' if model == 'F' else 'train_semisup.py'
    out = {}
    with rr.Reference() as R:
        tf = R.tf
        FLAGS = R.flags(use_one_hot=True, use_one_hot_boxpc=False, NUM_CHANNELS=6, restore_model_path=None, init_model_path=None,
                        init_class_ag_path=None, init_boxpc_path=None, SEMI_MODEL=model, BOX_PC_MASK_REPRESENTATION='A', **flags)
        FLAGS.TRAIN_CLS, FLAGS.TEST_CLS = FLAGS.SUNRGBD_SEMI_TRAIN_CLS, FLAGS.SUNRGBD_SEMI_TEST_CLS      # train_semisup_adv.py:78-79
        st = R.reset(v, requires_grad=True, dropout_masks=masks, feeds=[None if k is None else feed[k] for k in SEMI_FEED_ORDER] + [True])
        R.quiet()
        ns = rr.exec_train_graph(R, script, dict(
            FLAGS=FLAGS, MODEL=R.mod('semisup_v1_sunrgbd'), tf_util=R.mod('tf_util'), weak_losses=R.mod('weak_losses'), BATCH_SIZE=B,
            NUM_POINT=N, GPU_INDEX=0, BASE_LEARNING_RATE=0.001, BASE_LEARNING_RATE_D=0.001, DECAY_STEP=800000, DECAY_RATE=0.5,
            OPTIMIZER='adam', OPTIMIZER_D='sgd', MOMENTUM=0.9, BN_DECAY_DECAY_STEP=800000.))
        R.quiet(False)
        op = ns['train_semi_op']
        assert op.loss is ns['semi_loss']
        tv = op.var_list
        if tv is None:          # train_semisup.py:250: minimize over every trainable variable
            tv = [x for x in tf.get_collection(tf.GraphKeys.TRAINABLE_VARIABLES) if x.t.is_floating_point()]
        grads = torch.autograd.grad(ns['semi_loss'].t, [x.t for x in tv], allow_unused=True)
        moving = {k: _np(var.t) for k, var in st.vars.items() if 'moving' in k}
        _pack_step(out, ns['semi_loss'].t.detach(), {x.name[:-2]: _np(g) for x, g in zip(tv, grads)}, moving)
        pack(out, 'logits', ns['logits'])
        for k in (SEMI_EP_KEYS_F if model == 'F' else SEMI_EP_KEYS_A):
            pack(out, 'ep.' + k, ns['end_points'][k])
    return out


def _semi_train_oracle(model, flags):
    v, feed, masks = _semi_inputs(model)
    if model == 'F':
        from oracle import train_semisup_adv as ot
        FLAGS = config.cfg(**flags)
    else:
        from oracle import train_semisup as ot
        FLAGS = config.cfg(SEMI_MODEL='A', **flags)
    loss, grads, vs, ep = ot.loss_and_grads(v, FLAGS, feed, masks, global_step=0, dtype=F64)
    out = {}
    moving = {k: _np(t) for k, t in vs.vars.items() if 'moving' in k}
    _pack_step(out, loss, {k: _np(g) for k, g in grads.items()}, moving)
    pack(out, 'logits', ep['logits'])
    for k in (SEMI_EP_KEYS_F if model == 'F' else SEMI_EP_KEYS_A):
        pack(out, 'ep.' + k, ep[k])
    return out


def _case(ref, orc, *args, **kw):
    lean = kw.get('lean', False)

    def run(fn):
        def thunk():
            LEAN[0] = lean
            try:
                out = fn(*args)
            finally:
                LEAN[0] = False
            if lean:
                out = {k: v for k, v in out.items() if not k.startswith(('ep.', 'logits', 'moving.'))}
            return out
        return thunk
    return run(ref), run(orc)


REPROJ_VARIANTS = [
    dict(WEAK_REPROJECTION_CLIP_PRED_BOX=True),
    dict(WEAK_REPROJECTION_CLIP_LOWERB_LOSS=False),
    dict(WEAK_REPROJECTION_LOSS_TYPE='mse', WEAK_DIMS_LOSS_TYPE='mse', WEAK_REPROJECTION_ONLY_ON_2D_CLS=False),
    dict(WEAK_REPROJECTION_USE_SOFTMAX_PROJ=True, WEAK_TRAIN_BOX_W_REPROJECTION=[True, False, True]),
    dict(SEMI_BOXPC_FIT_ONLY_ON_2D_CLS=True, SEMI_INTRACLSDIMS_ONLY_ON_2D_CLS=False, WEAK_WEIGHT_REPROJECTION=1.0),
    dict(WEAK_WEIGHT_INACTIVE_VOLUME=1.5, WEAK_INACTIVE_VOL_ONLY_ON_2D_CLS=False,
         WEAK_INACTIVE_VOL_LOSS_MARGINS=[4.0, 1.5, 2.0, 0.4, 0.3, 1.0, 0.8, 0.3, 1.0, 0.6]),
    dict(WEAK_WEIGHT_SURFACE=2.0, WEAK_TRAIN_SEG_W_SURFACE=True, WEAK_TRAIN_BOX_W_SURFACE=[True, True, False], WEAK_SURFACE_MARGIN=0.05),
    dict(SEMI_REFINE_USING_BOXPC_DELTA_NUM=2, SEMI_WEIGH_BOXPC_DELTA_DURING_TEST=True, SEMI_BOXPC_MIN_FIT_LOSS_AFT_REFINE=False),
]
BOXPC_VARIANTS = [
    dict(BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF=True, BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA=True),
    dict(BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF=True, BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA=False),
    dict(BOXPC_WEIGH_DELTA_LOSS_BY_CLS_GT=True),
    dict(BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF=True, BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA=True),
    dict(BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF=True, BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA=False),
    dict(BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF=True, BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF=True, BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA=False,
         BOXPC_DELTA_LOSS_TYPE='mse'),
    dict(NORMALIZE_PC_BEFORE_SEG=True, NORMALIZATION_METHOD='01'),
    dict(NORMALIZE_PC_BEFORE_SEG=True, NORMALIZATION_METHOD='mean_zero_unit_var'),
]


CASES = {
    # name: (reference thunk, oracle thunk)
    'model_F_test_graph_refine2': _case(_model_F_reference, _model_F_oracle, 2, False),
    'model_F_test_graph_masked_pc': _case(_model_F_reference, _model_F_oracle, 1, True),
    'boxpc_train_rep_A': _case(_boxpc_train_reference, _boxpc_train_oracle, 'A', dict(BOXPC_WEIGHT_DELTA=4.)),
    'boxpc_train_rep_B': _case(_boxpc_train_reference, _boxpc_train_oracle, 'B', dict(BOXPC_WEIGHT_DELTA=4.)),
    'semisup_adv_train_cfg5': _case(_semi_train_reference, _semi_train_oracle, 'F', CFG5),
    'semisup_adv_train_tnet_only': _case(_semi_train_reference, _semi_train_oracle, 'F', dict(CFG5, SEMI_TRAIN_BOX_TRAIN_CLASS_AG_BOX=False)),
    'semisup_A_train': _case(_semi_train_reference, _semi_train_oracle, 'A', CFG_A),
    'semisup_A_train_softmax_proj': _case(_semi_train_reference, _semi_train_oracle, 'A',
                                          dict(CFG_A, WEAK_TRAIN_BOX_W_SURFACE=[True, False, True], WEAK_REPROJECTION_USE_SOFTMAX_PROJ=True), lean=True),
}
for _i, _f in enumerate(REPROJ_VARIANTS):
    CASES['semisup_adv_train_variant%d' % _i] = _case(_semi_train_reference, _semi_train_oracle, 'F', dict(CFG5, **_f), lean=True)
for _i, _f in enumerate(BOXPC_VARIANTS):
    CASES['boxpc_train_variant%d' % _i] = _case(_boxpc_train_reference, _boxpc_train_oracle, 'A' if _i != 7 else 'B',
                                                dict(BOXPC_WEIGHT_DELTA=4., **_f), lean=True)


def fixture_path(name):
    return os.path.join(HERE, 'ref_%s.npz' % name)


def compare(got, want, rtol=1e-9, atol=1e-10):
    """Every key of `want` (the reference's outputs) present in `got` with |got - want| <= atol + rtol * scale (scale: the tensor's largest magnitude; for a
    signature, its norm entry).  Returns the list of failures (empty: identical)."""
    bad = []
    for k in sorted(want):
        if k not in got:
            bad.append((k, 'missing'))
            continue
        a, b = np.asarray(got[k], dtype=np.float64), np.asarray(want[k], dtype=np.float64)
        if a.shape != b.shape:
            bad.append((k, 'shape %s vs %s' % (a.shape, b.shape)))
            continue
        if not a.size:
            continue
        scale = float(np.abs(b).max()) if not k.endswith('#sig') else float(b[0])
        err = float(np.abs(a - b).max())
        if not err <= atol + rtol * max(scale, 1.0):
            bad.append((k, 'max err %.3g (scale %.3g)' % (err, scale)))
    return bad
