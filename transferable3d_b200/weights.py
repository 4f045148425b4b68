"""Network tables and deterministic synthetic weights keyed by TF variable names.

The variable names are the reference's checkpoint contract (models/tf_util.py:1300-1318,
1484-1491; tf.contrib.layers.batch_norm default names):
    <scope>/<layer>/weights, /biases, /bn/beta, /bn/gamma, /bn/moving_mean, /bn/moving_variance
Conv kernels are [1, kw, Cin, Cout] (first layers use kw=D on a (B,N,D,1) image,
sunrgbd_detection/semisup_models.py:76,354), FC weights are [Cin, Cout].

Layer tables follow semisup_models.py:76-135 (inst_seg), :172-198 (tnet), :224-261 (box_est),
:354-392 (box_pc_mask_model rep A), semisup_v1_sunrgbd.py:188-197 (box_refine).
"""
import numpy as np

from .constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER, NUM_CLASS, BN_EPS

BOX_OUT = 3 + NUM_HEADING_BIN * 2 + NUM_SIZE_CLUSTER * 4   # 67
BOXPC_OUT = 2 + 3 + 3 + 1                                   # 9


def net_table(net, num_channel=6, one_hot=False, norm_box2d=False):
    """Returns [(layer_name, kind, kw, cin, cout, bn)] for one sub-network.

    kind: 'conv' ([1,kw,cin,cout] kernel) or 'fc' ([cin,cout]).
    """
    oh = NUM_CLASS if one_hot else 0
    nb = 4 if norm_box2d else 0
    if net == 'inst_seg':
        return [('conv1', 'conv', num_channel, 1, 64, True),
                ('conv2', 'conv', 1, 64, 64, True),
                ('conv3', 'conv', 1, 64, 64, True),
                ('conv4', 'conv', 1, 64, 128, True),
                ('conv5', 'conv', 1, 128, 1024, True),
                ('conv6', 'conv', 1, 64 + 1024 + oh, 512, True),
                ('conv7', 'conv', 1, 512, 256, True),
                ('conv8', 'conv', 1, 256, 128, True),
                ('conv9', 'conv', 1, 128, 128, True),
                ('conv10', 'conv', 1, 128, 2, False)]
    if net == 'tnet':
        return [('conv-reg1-stage1', 'conv', 1, 3, 128, True),
                ('conv-reg2-stage1', 'conv', 1, 128, 128, True),
                ('conv-reg3-stage1', 'conv', 1, 128, 256, True),
                ('fc1-stage1', 'fc', 0, 256 + oh + nb, 256, True),
                ('fc2-stage1', 'fc', 0, 256, 128, True),
                ('fc3-stage1', 'fc', 0, 128, 3, False)]
    if net == 'box_est':
        return [('conv-reg1', 'conv', 1, 3, 128, True),
                ('conv-reg2', 'conv', 1, 128, 128, True),
                ('conv-reg3', 'conv', 1, 128, 256, True),
                ('conv-reg4', 'conv', 1, 256, 512, True),
                ('fc1', 'fc', 0, 512 + oh + nb, 512, True),
                ('fc2', 'fc', 0, 512, 256, True),
                ('fc3', 'fc', 0, 256, BOX_OUT, False)]
    if net == 'box_refine':
        return [('fc0', 'fc', 0, 512 + oh, 512, True),
                ('fc1', 'fc', 0, 512, 256, True),
                ('fc2', 'fc', 0, 256, BOX_OUT, False)]
    if net == 'box_pc_mask_model':
        return [('conv-reg1', 'conv', num_channel + 6, 1, 128, True),
                ('conv-reg2', 'conv', 1, 128, 128, True),
                ('conv-reg3', 'conv', 1, 128, 256, True),
                ('conv-reg4', 'conv', 1, 256, 512, True),
                ('fc1', 'fc', 0, 512 + oh, 512, True),
                ('fc2', 'fc', 0, 512, 256, True),
                ('fc3', 'fc', 0, 256, BOXPC_OUT, False)]
    if net == 'box_pc_mask_model_B':        # representation B (semisup_models.py:400-471): box FC stack + point conv stack
        return [('extract_box_feats/fc0', 'fc', 0, 7, 128, True),
                ('extract_box_feats/fc1', 'fc', 0, 128, 128, True),
                ('extract_box_feats/fc2', 'fc', 0, 128, 256, True),
                ('extract_box_feats/fc3', 'fc', 0, 256, 512, False),
                ('conv-reg1', 'conv', num_channel, 1, 128, True),
                ('conv-reg2', 'conv', 1, 128, 128, True),
                ('conv-reg3', 'conv', 1, 128, 256, True),
                ('conv-reg4', 'conv', 1, 256, 512, True),
                ('fc1', 'fc', 0, 1024 + oh, 512, True),
                ('fc2', 'fc', 0, 512, 512, True),
                ('fc3', 'fc', 0, 512, 256, True),
                ('fc4', 'fc', 0, 256, BOXPC_OUT, False)]
    raise KeyError(net)


def _xavier_uniform(rng, shape):
    # tf.contrib.layers.xavier_initializer (models/tf_util.py:1183): uniform,
    # limit sqrt(6/(fan_in+fan_out)) with receptive field folded into the fans.
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:
        rf = shape[0] * shape[1]
        fan_in, fan_out = rf * shape[2], rf * shape[3]
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def init_net(rng, scope, table, out=None):
    """Synthetic variables for one sub-network (SURVEY 8d 'Synthetic weights')."""
    out = {} if out is None else out
    for name, kind, kw, cin, cout, bn in table:
        p = '%s/%s' % (scope, name)
        shape = (1, kw, cin, cout) if kind == 'conv' else (cin, cout)
        out[p + '/weights'] = _xavier_uniform(rng, shape)
        out[p + '/biases'] = rng.normal(0, 0.05, size=cout).astype(np.float32)
        if bn:
            out[p + '/bn/gamma'] = rng.uniform(0.8, 1.2, size=cout).astype(np.float32)
            out[p + '/bn/beta'] = rng.normal(0, 0.1, size=cout).astype(np.float32)
            out[p + '/bn/moving_mean'] = rng.normal(0, 0.1, size=cout).astype(np.float32)
            out[p + '/bn/moving_variance'] = rng.uniform(0.5, 1.5, size=cout).astype(np.float32)
    return out


def make_weights_model_F(seed=42, num_channel=6, use_one_hot=True, norm_box2d=False,
                         with_boxpc=True, boxpc_one_hot=False):
    """Variables of the model-F test/adv graph (test_semisup.py:61-149):
    class_agnostic/{inst_seg,tnet,box_est}, class_dependent/box_refine and
    D_boxpc_branch/box_pc_mask_model."""
    rng = np.random.Generator(np.random.PCG64(seed))
    v = {}
    init_net(rng, 'class_agnostic/inst_seg', net_table('inst_seg', num_channel), v)
    init_net(rng, 'class_agnostic/tnet', net_table('tnet', norm_box2d=norm_box2d), v)
    init_net(rng, 'class_agnostic/box_est', net_table('box_est', norm_box2d=norm_box2d), v)
    init_net(rng, 'class_dependent/box_refine', net_table('box_refine', one_hot=use_one_hot), v)
    if with_boxpc:
        init_net(rng, 'D_boxpc_branch/box_pc_mask_model',
                 net_table('box_pc_mask_model', num_channel, one_hot=boxpc_one_hot), v)
    return v


def make_weights_model_A(seed=42, num_channel=6, use_one_hot=True, norm_box2d=False):
    """Variables of model A / the F-PointNet v1 pipeline (semisup_v1_sunrgbd.py:81-130):
    inst_seg, tnet, box_est with the one-hot routed into all three."""
    rng = np.random.Generator(np.random.PCG64(seed))
    v = {}
    init_net(rng, 'inst_seg', net_table('inst_seg', num_channel, one_hot=use_one_hot), v)
    init_net(rng, 'tnet', net_table('tnet', one_hot=use_one_hot, norm_box2d=norm_box2d), v)
    init_net(rng, 'box_est', net_table('box_est', one_hot=use_one_hot, norm_box2d=norm_box2d), v)
    return v


def make_weights_boxpc(seed=43, num_channel=6, use_one_hot=False, rep='A', scope='box_pc_mask_model'):
    """Variables of the standalone BoxPC graph (train_boxpc.py:252); rep = BOX_PC_MASK_REPRESENTATION ('A' or 'B')."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return init_net(rng, scope, net_table('box_pc_mask_model' if rep == 'A' else 'box_pc_mask_model_B', num_channel,
                                          one_hot=use_one_hot))


def apply_logit_margin(variables, scope, k):
    """'random+margin(k)' regime of SURVEY App. D: scale conv10 by k so the mask logits are
    bimodal like a trained net.  Must be reported next to any mask-exactness figure."""
    v = dict(variables)
    v[scope + '/conv10/weights'] = variables[scope + '/conv10/weights'] * np.float32(k)
    v[scope + '/conv10/biases'] = variables[scope + '/conv10/biases'] * np.float32(k)
    return v


def fold_bn(variables, layer):
    """Eval-mode BN folded into (W[Cin_total, Cout], b[Cout]) in float64 then cast to f32.

    y = gamma*(xW+b-mu)/sqrt(var+eps)+beta  (models/tf_util.py:1308-1322,1660)
    """
    w = variables[layer + '/weights'].astype(np.float64)
    if w.ndim == 4:
        w = w.reshape(-1, w.shape[-1])          # [1,kw,cin,cout] -> [kw*cin, cout]
    b = variables[layer + '/biases'].astype(np.float64)
    if (layer + '/bn/gamma') in variables:
        g = variables[layer + '/bn/gamma'].astype(np.float64)
        beta = variables[layer + '/bn/beta'].astype(np.float64)
        mu = variables[layer + '/bn/moving_mean'].astype(np.float64)
        var = variables[layer + '/bn/moving_variance'].astype(np.float64)
        s = g / np.sqrt(var + BN_EPS)
        w = w * s[None, :]
        b = (b - mu) * s + beta
    return np.ascontiguousarray(w.astype(np.float32)), np.ascontiguousarray(b.astype(np.float32))


def save_npz(path, variables):
    np.savez(path, **{k.replace('/', '|'): v for k, v in variables.items()})


def load_npz(path):
    z = np.load(path)
    return {k.replace('|', '/'): z[k] for k in z.files}


def _np_layer(x, variables, layer, relu=True):
    w, b = fold_bn(variables, layer)
    y = x @ w.astype(np.float64) + b.astype(np.float64)
    return np.maximum(y, 0) if relu else y


def calibrate_seg_logits(variables, scope, pc, one_hot=None, target_frac=0.4, margin_std=2.0):
    """Synthetic-weight tooling (not on the product path): random Xavier weights give mask
    margins l1-l0 with a large common offset and tiny per-point spread (every point masked
    out).  This rescales conv10 so that, on the calibration frustums `pc`, the margin has
    standard deviation `margin_std` and a fraction `target_frac` of points is masked in --
    the 'random+margin(k)' regime of SURVEY App. D.  Returns (new variables, k, shift);
    k must be reported next to any mask-exactness figure.  Plain float64 numpy forward of
    the eval-mode seg net with BN folded (semisup_models.py:76-135)."""
    x = pc.astype(np.float64)
    B, N, _ = x.shape
    a = _np_layer(x, variables, scope + '/conv1')
    a = _np_layer(a, variables, scope + '/conv2')
    pf = _np_layer(a, variables, scope + '/conv3')
    a = _np_layer(pf, variables, scope + '/conv4')
    a = _np_layer(a, variables, scope + '/conv5')
    g = a.max(axis=1)
    if one_hot is not None:
        g = np.concatenate([g, one_hot.astype(np.float64)], axis=1)
    w6, b6 = fold_bn(variables, scope + '/conv6')
    a = np.maximum(pf @ w6[:64].astype(np.float64) + (g @ w6[64:].astype(np.float64))[:, None, :] + b6, 0)
    a = _np_layer(a, variables, scope + '/conv7')
    a = _np_layer(a, variables, scope + '/conv8')
    a = _np_layer(a, variables, scope + '/conv9')
    w10 = variables[scope + '/conv10/weights'].reshape(128, 2).astype(np.float64)
    b10 = variables[scope + '/conv10/biases'].astype(np.float64)
    d = a @ (w10[:, 1] - w10[:, 0]) + (b10[1] - b10[0])
    k = margin_std / max(d.std(), 1e-12)
    thr = np.quantile(d, 1.0 - target_frac)
    v = dict(variables)
    v[scope + '/conv10/weights'] = (variables[scope + '/conv10/weights'].astype(np.float64) * k).astype(np.float32)
    nb = b10 * k
    nb[1] -= k * thr          # margin' = k*(d - thr)
    v[scope + '/conv10/biases'] = nb.astype(np.float32)
    return v, float(k), float(-k * thr)


def standard_model_F(seed=42, num_channel=6, calib_frustums=4, calib_seed=999, margin_std=2.0):
    """The synthetic model-F weight set used by tests and bench: Xavier weights (seed) with the
    seg logits calibrated on `calib_frustums` synthetic frustums (target 40 % masked-in,
    margin std 2.0)."""
    from . import synth
    v = make_weights_model_F(seed, num_channel)
    pc = synth.make_batch(calib_frustums, 2048, num_channel, seed=calib_seed)['pc']
    v, k, shift = calibrate_seg_logits(v, 'class_agnostic/inst_seg', pc, margin_std=margin_std)
    return v, {'margin_k': k, 'margin_shift': shift, 'target_frac': 0.4, 'margin_std': margin_std}


def standard_model_A(seed=42, num_channel=6, use_one_hot=True, calib_frustums=4, calib_seed=999):
    """Same for model A / the F-PointNet v1 pipeline (one-hot routed into seg/tnet/box)."""
    from . import synth
    v = make_weights_model_A(seed, num_channel, use_one_hot)
    b = synth.make_batch(calib_frustums, 2048, num_channel, seed=calib_seed)
    v, k, shift = calibrate_seg_logits(v, 'inst_seg', b['pc'], b['one_hot'] if use_one_hot else None)
    return v, {'margin_k': k, 'margin_shift': shift, 'target_frac': 0.4, 'margin_std': 2.0}
