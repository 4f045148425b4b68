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
