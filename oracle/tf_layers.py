"""Layer wrappers: restatement of models/tf_util.py:1258-1323 (conv2d), :1463-1499
(fully_connected), :1501-1524 (max_pool2d), :1645-1664 (batch_norm_template), :1720-1741
(dropout).  Order is x.W + b -> BN -> activation, exactly as the reference.
"""
import contextlib
import numpy as np
import torch

BN_EPS = 1e-3       # tf.contrib.layers.batch_norm default epsilon (tf_util.py:1660)


class VarStore(object):
    """Plays the role of the TF variable collection + tf.variable_scope: variables are looked
    up by their full TF name.  BN moving statistics are updated in place in training mode
    (updates_collections=None, tf_util.py:1662)."""

    def __init__(self, variables, dtype=torch.float32, requires_grad=False):
        self.dtype = dtype
        self.vars = {}
        for k, v in variables.items():
            t = torch.as_tensor(np.asarray(v)).to(dtype).clone()
            if requires_grad and not k.endswith(('moving_mean', 'moving_variance')):
                t.requires_grad_(True)
            self.vars[k] = t
        self._scope = []
        self.dropout_masks = {}     # full scope name -> explicit keep mask (0/1 tensor)
        self.literal = True         # literal tiled-global conv6 (reference graph) vs folded

    @contextlib.contextmanager
    def variable_scope(self, name):
        self._scope.append(name)
        try:
            yield
        finally:
            self._scope.pop()

    def scope_name(self, name=None):
        parts = [s for s in self._scope if s]
        if name:
            parts.append(name)
        return '/'.join(parts)

    def get(self, name):
        return self.vars[self.scope_name(name)]

    def has(self, name):
        return self.scope_name(name) in self.vars


def _as_bool(is_training):
    return bool(is_training)


def batch_norm(x, vs, scope, is_training, bn_decay):
    """tf_util.py:1645-1664. x: (..., C). Train: batch mean / biased variance over all axes but
    the last; moving <- decay*moving + (1-decay)*batch, with the Bessel-corrected variance
    going into the moving average (TF1 fused-BN runtime behaviour)."""
    decay = 0.9 if bn_decay is None else float(bn_decay)
    with vs.variable_scope(scope):
        gamma, beta = vs.get('gamma'), vs.get('beta')
        mm, mv = vs.get('moving_mean'), vs.get('moving_variance')
        if _as_bool(is_training):
            red = tuple(range(x.dim() - 1))
            n = 1
            for d in red:
                n *= x.shape[d]
            mean = x.mean(dim=red)
            var = ((x - mean) ** 2).mean(dim=red)
            with torch.no_grad():
                mm.mul_(decay).add_((1 - decay) * mean.detach())
                unb = var.detach() * (float(n) / max(n - 1, 1))
                mv.mul_(decay).add_((1 - decay) * unb)
        else:
            mean, var = mm, mv
        return (x - mean) * (gamma / torch.sqrt(var + BN_EPS)) + beta


def _linear(x, vs, num_out, scope, bn, is_training, activation_fn, bn_decay):
    with vs.variable_scope(scope):
        w = vs.get('weights')
        w2 = w.reshape(-1, w.shape[-1])             # [1,kw,cin,cout] -> [kw*cin,cout]
        assert w2.shape[1] == num_out and w2.shape[0] == x.shape[-1], \
            (vs.scope_name(), tuple(w.shape), tuple(x.shape))
        y = x @ w2 + vs.get('biases')
        if bn:
            y = batch_norm(y, vs, 'bn', is_training, bn_decay)
        if activation_fn is not None:
            y = activation_fn(y)
        return y


def conv2d(x, num_out, kernel_size, vs, scope, bn=False, is_training=False,
           activation_fn=torch.relu, bn_decay=None):
    """Per-point 1x1 conv (or the [1,D] first-layer kernel on the (B,N,D,1) image, which is the
    same linear map over the D input channels).  x: (B,N,Cin)."""
    return _linear(x, vs, num_out, scope, bn, is_training, activation_fn, bn_decay)


def fully_connected(x, num_out, vs, scope, bn=False, is_training=False,
                    activation_fn=torch.relu, bn_decay=None):
    """x: (B,Cin) (tf_util.py:1463-1499)."""
    return _linear(x, vs, num_out, scope, bn, is_training, activation_fn, bn_decay)


def max_pool_points(x):
    """tf_util.max_pool2d(net, [num_point,1]) + squeeze: (B,N,C) -> (B,C)."""
    return x.max(dim=1).values


def dropout(x, vs, is_training, scope, keep_prob=0.5):
    """tf_util.py:1720-1741 / tf.nn.dropout: x*mask/keep when training, identity otherwise.
    The keep mask is an explicit input (vs.dropout_masks[full scope name])."""
    if not _as_bool(is_training):
        return x
    name = vs.scope_name(scope)
    if name not in vs.dropout_masks:
        raise KeyError('training-mode dropout needs an explicit keep mask for %s' % name)
    m = torch.as_tensor(vs.dropout_masks[name]).to(x.dtype)
    return x * m / keep_prob


def leaky_relu(x):
    return torch.nn.functional.leaky_relu(x, 0.2)     # tf.nn.leaky_relu default alpha
