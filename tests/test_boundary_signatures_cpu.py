"""The drop-in boundary, checked against the reference's own definitions: for every function that a reference module on the
path defines (tests/golden/ref_signatures.json, written by make_reference_golden.py from inspect.signature of the reference's
functions) and that the same-named module of transferable3d_b200 also defines, the reference's parameters must be a positional
PREFIX of the product's -- same names, same order, same defaults -- so that a caller written against the reference, positional or
keyword, works unchanged.  The product may append keyword arguments of its own (device=, FLAGS=, variables=, ...).
Deviations are listed here, one by one, with the reason."""
import importlib
import json
import os
import sys

import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
sys.path.insert(0, GOLDEN)
import reference_runner as rr  # noqa: E402

REF = json.load(open(os.path.join(GOLDEN, 'ref_signatures.json')))

# (module, function, parameter) -> (reference default, product default, why)
ALLOWED_DEFAULTS = {
    ('tf_util', 'conv2d', 'activation_fn'): ('<fn:relu>', 'relu', 'activations are named: there is no tf.nn.relu object on this side'),
    ('tf_util', 'fully_connected', 'activation_fn'): ('<fn:relu>', 'relu', 'same'),
    ('tf_util', 'max_pool2d', 'scope'): ('<required>', None, 'the scope only names a TF op; optional here'),
}


def _pairs():
    for mod, fns in sorted(REF.items()):
        try:
            pm = importlib.import_module('transferable3d_b200.' + mod)
        except ImportError:
            continue
        for name in sorted(fns):
            if name in rr.signatures_of(pm):
                yield mod, name


PAIRS = list(_pairs())


def test_every_path_module_is_mirrored():
    have = {m for m, _ in PAIRS}
    assert have == set(REF), sorted(set(REF) - have)
    assert len(PAIRS) >= 60, len(PAIRS)


@pytest.mark.parametrize('mod,name', PAIRS)
def test_reference_parameters_are_a_positional_prefix(mod, name):
    want = REF[mod][name]
    got = rr.signatures_of(importlib.import_module('transferable3d_b200.' + mod))[name]
    assert [p for p, _ in got[:len(want)]] == [p for p, _ in want], (want, got)
    for (p, rd), (_, pd) in zip(want, got):
        if rd == pd or (isinstance(rd, (list, tuple)) and list(rd) == list(pd if isinstance(pd, (list, tuple)) else [])):
            continue
        allowed = ALLOWED_DEFAULTS.get((mod, name, p))
        assert allowed is not None and allowed[0] == rd and allowed[1] == pd, (mod, name, p, rd, pd)
    for p, d in got[len(want):]:
        assert d != '<required>', 'product-only parameter %s of %s.%s must be optional' % (p, mod, name)


@pytest.mark.skipif(not rr.available(), reason='reference tree not present')
def test_signature_fixture_is_live():
    assert json.loads(json.dumps(rr.reference_signatures())) == REF


def _error_cases(M, sm, cfg, T, tf_const_bool=None):
    """The reference's own `raise` sites on the path (semisup_v1_sunrgbd.py:79, :96, :254; semisup_models.py:323, :343, :421),
    each reached before any variable or kernel is touched.  M / sm: the two modules of one side; cfg(**kw): its flags; T: tensor
    constructor.  -> [(exception type name, message)]."""
    pc, oh = T((3, 8, 6)), T((3, 10))          # batch >= 3: tf_get_box_pc_representation reads dims_reg[2] (tf_util.py:780, unused)
    box = (T((3, 3)), T((3, 3)), T((3,)))
    f = False if tf_const_bool is None else tf_const_bool
    calls = [
        lambda: M.get_semi_model(pc, None, None, oh, f, True, c=cfg(SEMI_MODEL='Q')),
        lambda: M.get_semi_loss((None, None), (None,) * 14, {}, c=cfg(SEMI_MODEL='Q')),
        lambda: M.get_semi_model_backbone(pc, None, None, oh, f, True, oracle_mask=T((3, 8)), c=cfg(SEMI_MODEL='A')),
        lambda: sm.box_pc_mask_features_model(box, pc, None, 9, f, {}, None, False, c=cfg(BOX_PC_MASK_REPRESENTATION='Z')),
        lambda: sm.box_pc_mask_features_model(box, pc, None, 9, f, {}, None, False, normalize_pc=True, normalize_method='bad',
                                              c=cfg(BOX_PC_MASK_REPRESENTATION='A')),
    ]
    out = []
    for call in calls:
        try:
            call()
            out.append(('no exception', ''))
        except Exception as e:          # noqa: BLE001
            out.append((type(e).__name__, str(e)))
    return out


EXPECTED_ERRORS = [('Exception', 'Not implemented SEMI_MODEL: Q'), ('Exception', 'Not implemented SEMI_MODEL: Q'), ('NotImplementedError', ''),
                   ('Exception', 'Box pc mask representation not implemented: Z'), ('Exception', 'Invalid normalization method')]


def test_error_behaviour_matches_the_reference_raise_sites():
    import torch
    from transferable3d_b200 import semisup_v1_sunrgbd as M, semisup_models as sm, config
    assert _error_cases(M, sm, config.cfg, lambda shape: torch.zeros(shape)) == EXPECTED_ERRORS


@pytest.mark.skipif(not rr.available(), reason='reference tree not present')
def test_expected_errors_are_what_the_reference_raises():
    import numpy as np
    with rr.Reference() as R:
        tf = R.tf
        R.reset({})
        R.quiet()
        try:
            got = _error_cases(R.mod('semisup_v1_sunrgbd'), R.mod('semisup_models'), lambda **kw: R.flags(**kw),
                               lambda shape: tf.constant(np.zeros(shape), dtype=tf.float32), tf.constant(False))
        finally:
            R.quiet(False)
    assert got == EXPECTED_ERRORS, got


# ---- the checkpoint contract: variable names and shapes -------------------------------------------------------------
VARIABLES = json.load(open(os.path.join(GOLDEN, 'ref_variables.json')))


def _product_tables():
    from transferable3d_b200 import weights
    return {'model_F_test_graph': weights.make_weights_model_F(seed=11),
            'model_F_test_graph_box2d_feats': weights.make_weights_model_F(seed=11, norm_box2d=True),
            'model_A_train_graph': weights.make_weights_model_A(),
            'boxpc_rep_A_train_graph': weights.make_weights_boxpc(rep='A'),
            'boxpc_rep_B_train_graph': weights.make_weights_boxpc(rep='B')}


@pytest.mark.parametrize('graph', sorted(VARIABLES))
def test_variable_names_and_shapes_are_the_ones_the_reference_graph_creates(graph):
    """Every tf.get_variable(name, shape) the reference's graph code issues (recorded while its own get_model / train() graphs
    were built on the stand-in) against the weight tables of transferable3d_b200.weights: the same set of names, the same
    shapes -- what a TensorFlow checkpoint of the reference holds is what the variable store expects, and nothing else."""
    import numpy as np
    want = VARIABLES[graph]
    got = {k: list(np.shape(v)) for k, v in _product_tables()[graph].items()}
    assert sorted(got) == sorted(want), (sorted(set(got) ^ set(want))[:8])
    assert got == {k: list(v) for k, v in want.items()}
    assert len(want) >= 38


@pytest.mark.skipif(not rr.available(), reason='reference tree not present')
def test_variable_fixture_is_live():
    assert json.loads(json.dumps(rr.reference_variables())) == VARIABLES


# ---- the restore maps of the pretrained branches ----------------------------------------------------------------------------------------
RESTORE_MAPS = json.load(open(os.path.join(GOLDEN, 'ref_restore_maps.json')))


def test_remap_scope_reproduces_the_reference_restore_maps():
    """load_variable_scopes_from_ckpt (train_semisup_adv.py:224-237), run as is on the cfg5 graph with the scope lists of :450-467,
    hands tf.train.Saver {name in the pretrained checkpoint: graph variable}: tf_checkpoint.remap_scope must send exactly those
    checkpoint names (a model-A checkpoint, a BoxPC checkpoint) to exactly those variables, dropping the optimizer slots."""
    from transferable3d_b200 import tf_checkpoint as ck, weights
    tables = {'class_agnostic': weights.make_weights_model_A(), 'D_boxpc_branch': weights.make_weights_boxpc(rep='A')}
    for scope, m in RESTORE_MAPS.items():
        assert sorted(m) == sorted(tables[scope])                          # what such a checkpoint holds
        ckpt = dict(tables[scope])
        ckpt.update({'beta1_power': 0.9, 'beta2_power': 0.999, sorted(m)[0] + '/Adam': 0, sorted(m)[0] + '/Adam_1': 0})
        got = ck.remap_scope(ckpt, scope + '/')
        assert sorted(got) == sorted(m.values())
        assert all(got[v] is ckpt[k] for k, v in m.items())
    assert len(RESTORE_MAPS['class_agnostic']) == 126 and len(RESTORE_MAPS['D_boxpc_branch']) == 38


@pytest.mark.skipif(not rr.available(), reason='reference tree not present')
def test_restore_map_fixture_is_live():
    assert json.loads(json.dumps(rr.reference_restore_maps())) == RESTORE_MAPS
