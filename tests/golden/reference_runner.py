"""Runs the reference's own Python files (from /root/reference, unmodified, where they lie) on the TF1 stand-in of
tests/golden/tf1_shim.py.  TEST INFRASTRUCTURE: imported by make_reference_golden.py (which writes the committed fixtures
tests/golden/ref_*.npz) and by tests/test_oracle_vs_reference_cpu.py for the live re-run when the reference is present.

Modules the reference imports but does not ship / that need Python 2:
  box_util   (frustum-pointnets train/box_util.py, absent from the reference tree): only `box3d_iou` is imported; it is
             provided by oracle/box_util.py (a restatement of the published polygon-clipping routine) and is not called by
             any pinned function except compute_box3d_iou (metrics, not on the compared outputs);
  cPickle    -> pickle (sunrgbd_data/utils.py:338).
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('T3D_REFERENCE_ROOT', '/root/reference')
REF_DIRS = ['models', 'sunrgbd/sunrgbd_detection', 'sunrgbd/sunrgbd_data']
REF_MODULES = ['tf_util', 'model_util', 'weak_losses', 'config', 'timer', 'utils', 'roi_seg_box3d_dataset', 'semisup_models',
               'semisup_v1_sunrgbd', 'boxpc_sunrgbd', 'box_pc_fit_dataset', 'eval_det', 'test_semisup', 'train_boxpc',
               'train_semisup', 'train_semisup_adv', 'train_util', 'roi_semi_dataset', 'box_util', 'cPickle', 'tensorflow']


def available():
    return os.path.isdir(os.path.join(REF, 'models'))


class Reference(object):
    """Context manager: installs the shim + stubs, puts the reference directories on sys.path, imports on demand, and
    removes every trace on exit (the reference's module names -- tf_util, utils, config -- must not leak into pytest)."""

    def __enter__(self):
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        if HERE not in sys.path:
            sys.path.insert(0, HERE)
        self._saved_modules = {k: sys.modules.get(k) for k in REF_MODULES}
        for k in REF_MODULES:
            sys.modules.pop(k, None)
        self._saved_path = list(sys.path)
        self._saved_argv = list(sys.argv)
        sys.argv = [sys.argv[0]]
        import tf1_shim
        self.tf = tf1_shim.install()
        import pickle
        sys.modules['cPickle'] = pickle
        from oracle import box_util as obu
        bu = types.ModuleType('box_util')
        bu.box3d_iou = obu.box3d_iou
        sys.modules['box_util'] = bu
        for d in REF_DIRS:
            sys.path.insert(0, os.path.join(REF, d))
        self._stdout = sys.stdout
        return self

    def __exit__(self, *exc):
        sys.stdout = self._stdout
        sys.path[:] = self._saved_path
        sys.argv = self._saved_argv
        for k in REF_MODULES:
            sys.modules.pop(k, None)
            if self._saved_modules[k] is not None:
                sys.modules[k] = self._saved_modules[k]
        return False

    def quiet(self, on=True):
        """The reference prints every layer's shape while building a graph."""
        sys.stdout = open(os.devnull, 'w') if on else self._stdout

    def mod(self, name):
        m = importlib.import_module(name)
        f = getattr(m, '__file__', '') or ''
        if name not in ('box_util', 'cPickle', 'tensorflow'):
            assert os.path.realpath(f).startswith(os.path.realpath(REF)), (name, f)
        return m

    def flags(self, **overrides):
        """The reference's own defaults (models/config.py), as parse_special_args() returns them with no command line."""
        cfg = self.mod('config').cfg
        flags = cfg.parse_special_args()
        for k, v in overrides.items():
            setattr(flags, k, v)
        return flags

    def reset(self, variables, float_dtype=torch.float64, requires_grad=False, dropout_masks=None, feeds=None):
        import tf1_shim
        tf1_shim.set_float(float_dtype)
        tf1_shim.STATE.reset(variables, trainable_grad=requires_grad)
        tf1_shim.STATE.dropout_masks = dict(dropout_masks or {})
        tf1_shim.STATE.feeds = list(feeds or [])
        return tf1_shim.STATE


def to_np(v):
    """T / torch / nested containers -> numpy (float64 where floating)."""
    import tf1_shim
    if isinstance(v, tf1_shim.T):
        v = v.t
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy()
    if isinstance(v, (list, tuple)):
        return [to_np(x) for x in v]
    if isinstance(v, dict):
        return {k: to_np(x) for k, x in v.items()}
    return v


def _is_literal(node):
    import ast
    try:
        ast.literal_eval(node)
        return True
    except Exception:
        return False


def exec_train_graph(R, script, namespace):
    """Executes the graph-building statements of a reference training script -- the body of
    `with tf.Graph().as_default(): with tf.device(...):` inside its train() -- verbatim from the script's AST, in `namespace`.
    The script's module level cannot be imported (it parses a command line, creates log directories, copies files and loads the
    frustum pickles), so `namespace` supplies what that level would have bound (FLAGS, MODEL, BATCH_SIZE, NUM_POINT, ...);
    the script's own top-level functions (get_bn_decay, get_learning_rate, get_scope_vars, ...) and literal constants
    (BN_INIT_DECAY, ...) are taken from the same AST.  Returns the namespace with every local of the block bound."""
    import ast
    path = os.path.join(REF, 'sunrgbd/sunrgbd_detection', script)
    tree = ast.parse(open(path).read(), filename=path)
    pre = []
    train_fn = None
    for node in tree.body:
        if isinstance(node, ast.FunctionDef):
            if node.name == 'train':
                train_fn = node
            pre.append(node)
        elif isinstance(node, ast.Assign) and _is_literal(node.value) and all(isinstance(t, ast.Name) for t in node.targets):
            if not all(t.id in namespace for t in node.targets):
                pre.append(node)
    assert train_fn is not None, script
    block = None
    for node in train_fn.body:
        if isinstance(node, ast.With) and 'Graph' in ast.dump(node.items[0]):
            inner = node.body[0]
            assert isinstance(inner, ast.With) and 'device' in ast.dump(inner.items[0]), ast.dump(inner.items[0])
            block = inner.body
    assert block is not None, script
    ns = dict(namespace)
    ns.setdefault('tf', R.tf)
    ns.setdefault('np', np)
    ns.setdefault('log_string', lambda s: None)
    exec(compile(ast.Module(body=pre, type_ignores=[]), path, 'exec'), ns)
    ns['log_string'] = namespace.get('log_string', lambda s: None)
    exec(compile(ast.Module(body=block, type_ignores=[]), path, 'exec'), ns)
    ns['_block_lines'] = (block[0].lineno, max(getattr(n, 'end_lineno', n.lineno) for n in block))
    return ns


SIGNATURE_MODULES = ['semisup_v1_sunrgbd', 'boxpc_sunrgbd', 'semisup_models', 'model_util', 'weak_losses', 'tf_util', 'roi_seg_box3d_dataset',
                     'eval_det', 'test_semisup']


def encode_default(d):
    import inspect
    import json
    if d is inspect.Parameter.empty:
        return '<required>'
    if callable(d):
        return '<fn:%s>' % getattr(d, '__name__', str(d))
    try:
        json.dumps(d)
        return d
    except Exception:
        return repr(d)


def signatures_of(mod):
    """{function name: [[parameter, default], ...]} of the functions a module defines itself."""
    import inspect
    out = {}
    for name, fn in vars(mod).items():
        if inspect.isfunction(fn) and fn.__module__ == mod.__name__ and not name.startswith('_'):
            out[name] = [[p.name, encode_default(p.default)] for p in inspect.signature(fn).parameters.values()]
    return out


def reference_signatures():
    with Reference() as R:
        R.reset({})
        return {m: signatures_of(R.mod(m)) for m in SIGNATURE_MODULES}


def reference_config_defaults():
    with Reference() as R:
        return {k: v for k, v in vars(R.flags()).items() if k != 'config_str'}


def reference_variables():
    """{graph: {TF variable name: shape}}: every variable the reference's graph code creates (tf.get_variable calls recorded by the
    stand-in) for the graphs on the path -- the checkpoint contract."""
    import reference_cases as rc
    import tf1_shim
    out = {}

    def grab(name, thunk):
        thunk()
        out[name] = {k: list(v) for k, v in sorted(tf1_shim.STATE.requested_shapes.items())}
    grab('model_F_test_graph', lambda: rc._model_F_reference(1, False))
    grab('model_F_test_graph_box2d_feats', lambda: rc._model_F_reference(1, False, False, True))
    grab('model_A_train_graph', lambda: rc._semi_train_reference('A', rc.CFG_A))
    grab('boxpc_rep_A_train_graph', lambda: rc._boxpc_train_reference('A', dict(BOXPC_WEIGHT_DELTA=4.)))
    grab('boxpc_rep_B_train_graph', lambda: rc._boxpc_train_reference('B', dict(BOXPC_WEIGHT_DELTA=4.)))
    return out


def reference_restore_maps():
    """{scope: {name in the pretrained checkpoint: variable of the scoped graph}} as the reference's own
    load_variable_scopes_from_ckpt builds them for tf.train.Saver(var_list=...) (train_semisup_adv.py:224-237, called at :450-467
    with ['class_agnostic'] / [''] and ['D_boxpc_branch'] / ['']), on the cfg5 training graph."""
    import reference_cases as rc
    import tf1_shim
    from transferable3d_b200 import weights
    v, feed, masks = rc._semi_inputs('F')
    B, N = feed['pc'].shape[:2]
    out = {}
    with Reference() as R:
        FLAGS = R.flags(use_one_hot=True, use_one_hot_boxpc=False, NUM_CHANNELS=6, restore_model_path=None, init_model_path=None,
                        init_class_ag_path=None, init_boxpc_path=None, SEMI_MODEL='F', BOX_PC_MASK_REPRESENTATION='A', **rc.CFG5)
        FLAGS.TRAIN_CLS, FLAGS.TEST_CLS = FLAGS.SUNRGBD_SEMI_TRAIN_CLS, FLAGS.SUNRGBD_SEMI_TEST_CLS
        R.reset(v, dropout_masks=masks, feeds=[None if k is None else feed[k] for k in rc.SEMI_FEED_ORDER] + [True])
        R.quiet()
        ns = exec_train_graph(R, 'train_semisup_adv.py', dict(
            FLAGS=FLAGS, MODEL=R.mod('semisup_v1_sunrgbd'), tf_util=R.mod('tf_util'), weak_losses=R.mod('weak_losses'), BATCH_SIZE=B,
            NUM_POINT=N, GPU_INDEX=0, BASE_LEARNING_RATE=0.001, BASE_LEARNING_RATE_D=0.001, DECAY_STEP=800000, DECAY_RATE=0.5,
            OPTIMIZER='adam', OPTIMIZER_D='sgd', MOMENTUM=0.9, BN_DECAY_DECAY_STEP=800000.))
        R.quiet(False)
        for scope in ('class_agnostic', 'D_boxpc_branch'):
            n0 = len(tf1_shim.STATE.collections.get('savers', []))
            ns['load_variable_scopes_from_ckpt']([scope], [''], None, 'pretrained.ckpt')
            saver = tf1_shim.STATE.collections['savers'][n0]
            out[scope] = {k: var.op.name for k, var in sorted(saver.var_list.items())}
    return out
