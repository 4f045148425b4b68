"""The oracle against the REFERENCE'S OWN SOURCE (not against itself): tests/golden/ref_<case>.npz hold the outputs of the
reference's unmodified Python files -- test_semisup.get_model called as is, the graph blocks of train_boxpc.train(),
train_semisup_adv.train() and train_semisup.train() executed verbatim from the scripts' ASTs, model_util / tf_util / the numpy
helpers called directly -- run on the TF1 stand-in tests/golden/tf1_shim.py in float64 (make_reference_golden.py, run in the
container that has /root/reference).  Here the oracle runs the same seeded inputs, weights, flags and dropout masks in float64
and must reproduce every stored value to 1e-9 of its scale: losses, end points, which variables train, every gradient (as norm /
sum / four random projections), the updated moving statistics and the schedules.  Where the reference tree is present the
fixtures are also re-derived live, so a stale or hand-edited fixture cannot pass.

What this pins: the oracle's structure against the reference's code.  What it cannot pin: TensorFlow's kernels -- the arithmetic
of each op comes from the stand-in (documented TF1 semantics on PyTorch-CPU)."""
import json
import os
import sys

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
sys.path.insert(0, GOLDEN)
import reference_cases as rc  # noqa: E402
import reference_runner as rr  # noqa: E402

CASE_NAMES = sorted(rc.CASES)


@pytest.mark.parametrize('name', CASE_NAMES)
def test_oracle_reproduces_reference_fixture(name):
    path = rc.fixture_path(name)
    assert os.path.exists(path), 'missing fixture %s (python tests/golden/make_reference_golden.py %s)' % (path, name)
    want = dict(np.load(path))
    assert 'loss' in want or 'logits' in want or len(want) >= 3, sorted(want)[:5]
    got = rc.CASES[name][1]()
    bad = rc.compare(got, want)
    assert not bad, (name, len(bad), bad[:8])


@pytest.mark.skipif(not rr.available(), reason='reference tree not present (fixtures were generated where it is)')
@pytest.mark.parametrize('name', CASE_NAMES)
def test_fixture_is_what_the_reference_computes(name):
    want = dict(np.load(rc.fixture_path(name)))
    got = rc.CASES[name][0]()
    assert sorted(got) == sorted(want), (sorted(set(got) ^ set(want))[:8])
    bad = rc.compare(got, want, rtol=1e-10, atol=1e-11)      # float64 round-off only (thread count changes the summation order)
    assert not bad, (name, bad[:8])


def test_no_unlisted_fixture():
    have = sorted(f[4:-4] for f in os.listdir(GOLDEN) if f.startswith('ref_') and f.endswith('.npz'))
    assert have == CASE_NAMES, sorted(set(have) ^ set(CASE_NAMES))


def _product_flags():
    from transferable3d_b200 import config
    return vars(config.cfg())


def test_config_defaults_match_reference_config_py():
    """Every flag of transferable3d_b200.config.cfg() against models/config.py's own defaults (fixture written from
    config.cfg.parse_special_args() with no command line).  The two the reference leaves empty and sets per script
    (SEMI_MODEL, BOX_PC_MASK_REPRESENTATION: scripts/*.sh) and the command-line arguments of test_semisup.py that are not in
    config.py are listed explicitly."""
    ref = json.load(open(os.path.join(GOLDEN, 'ref_config_defaults.json')))
    script_level = {'SEMI_MODEL': 'F', 'BOX_PC_MASK_REPRESENTATION': 'A'}
    argparse_only = {'refine', 'mask_pc_for_boxpc', 'use_one_hot', 'use_one_hot_boxpc'}
    mismatched = {}
    for k, v in _product_flags().items():
        if k in script_level:
            assert ref[k] == '' and v == script_level[k], (k, ref[k], v)
        elif k in argparse_only and k not in ref:
            continue
        elif k == 'TEST_CLS':          # train_semisup_adv.py:79: FLAGS.TEST_CLS = FLAGS.SUNRGBD_SEMI_TEST_CLS
            assert list(v) == ref['SUNRGBD_SEMI_TEST_CLS']
        else:
            assert k in ref, 'product flag %s does not exist in the reference config' % k
            rv = ref[k]
            pv = list(v) if isinstance(v, (tuple, list)) else v
            if pv != rv:
                mismatched[k] = (pv, rv)
    assert not mismatched, mismatched


@pytest.mark.skipif(not rr.available(), reason='reference tree not present')
def test_config_fixture_is_live():
    with rr.Reference() as R:
        flags = R.flags()
        live = {k: v for k, v in vars(flags).items() if k != 'config_str'}
    ref = json.load(open(os.path.join(GOLDEN, 'ref_config_defaults.json')))
    assert json.loads(json.dumps(live)) == ref


def test_product_host_helpers_reproduce_reference_fixture():
    """The label codecs and voc_ap are host numpy in the product too (transferable3d_b200.roi_seg_box3d_dataset / eval_det): held
    to the reference's outputs directly (the device-side counterparts are compared in the -m gpu tests)."""
    from transferable3d_b200 import roi_seg_box3d_dataset as D, eval_det as E
    want = dict(np.load(rc.fixture_path('numpy_helpers')))
    got = rc._numpy_side(D, E, None, None, host_scalars_only=True)
    assert sorted(got) == ['angle2class', 'class2angle', 'class2size', 'rotate_pc_along_y', 'size2class.cls', 'size2class.res', 'voc_ap']
    assert not rc.compare(got, {k: want[k] for k in got}, rtol=1e-12, atol=1e-13)
