"""Known answers for the TF1 stand-in (tests/golden/tf1_shim.py) that the reference fixtures are produced with: the examples of
TensorFlow 1.x's own API documentation and hand-computed values for every op whose semantics are not the obvious ones.  The
stand-in supplies the arithmetic of each tf.* op when the reference's source is executed, so it gets its own tests."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
import tf1_shim  # noqa: E402


@pytest.fixture()
def tf():
    tf1_shim.set_float(torch.float64)
    tf1_shim.STATE.reset({})
    mod = tf1_shim.install()
    yield mod
    sys.modules.pop('tensorflow', None)


def n(x):
    return x.t.detach().numpy()


def test_dynamic_partition_doc_example(tf):
    # tf.dynamic_partition docs: partitions = [0, 0, 1, 1, 0], data = [10, 20, 30, 40, 50] -> [10, 20, 50], [30, 40]
    parts = tf.dynamic_partition(tf.constant([10., 20., 30., 40., 50.]), tf.constant([0, 0, 1, 1, 0]), 2)
    assert n(parts[0]).tolist() == [10., 20., 50.] and n(parts[1]).tolist() == [30., 40.]
    rows = tf.dynamic_partition(tf.constant(np.arange(12.).reshape(4, 3)), tf.constant([2, 0, 2, 1]), 3)
    assert n(rows[2]).tolist() == [[0., 1., 2.], [6., 7., 8.]] and n(rows[0]).shape == (1, 3)


def test_one_hot_doc_examples(tf):
    # tf.one_hot docs: indices = [0, 2, -1, 1], depth = 3, on 5.0, off 0.0 -> the -1 row is all off
    got = n(tf.one_hot(tf.constant([0, 2, -1, 1]), 3, on_value=5.0, off_value=0.0, axis=-1))
    assert got.tolist() == [[5., 0., 0.], [0., 0., 5.], [0., 0., 0.], [0., 5., 0.]]
    got = n(tf.one_hot(tf.constant([[0, 2], [1, -1]]), 3, on_value=1.0, off_value=0.0, axis=-1))
    assert got.shape == (2, 2, 3) and got[1, 1].tolist() == [0., 0., 0.]


def test_where_tf1_row_select_and_indices(tf):
    x, y = tf.constant(np.arange(6.).reshape(3, 2)), tf.constant(-np.ones((3, 2)))
    # TF1: a rank-1 condition whose size matches the first dimension selects whole rows
    assert n(tf.where(tf.constant([True, False, True]), x, y)).tolist() == [[0., 1.], [-1., -1.], [4., 5.]]
    assert n(tf.where(x > 2.5, x, y)).tolist() == [[-1., -1.], [-1., 3.], [4., 5.]]
    # single argument: coordinates of the true elements (docs example)
    assert n(tf.where(tf.constant([[True, False], [False, True]]))).tolist() == [[0, 0], [1, 1]]
    with pytest.raises(AssertionError):            # TF1's where does not broadcast a (3, 1) condition
        tf.where(tf.constant([[True], [False], [True]]), x, y)


def test_huber_and_mse_losses(tf):
    # tf.losses.huber_loss: 0.5 e^2 if |e| <= d else 0.5 d^2 + d (|e| - d); default reduction = mean over non-zero weights
    lab, pred = tf.constant([0., 0., 0., 0.]), tf.constant([0.5, -1.0, 3.0, -0.2])
    none = n(tf.losses.huber_loss(lab, pred, reduction=tf.losses.Reduction.NONE))
    assert np.allclose(none, [0.125, 0.5, 0.5 + 2.0, 0.02])
    assert np.isclose(float(n(tf.losses.huber_loss(lab, pred))), none.mean())
    d2 = n(tf.losses.huber_loss(lab, pred, delta=2.0, reduction=tf.losses.Reduction.NONE))
    assert np.allclose(d2, [0.125, 0.5, 2.0 + 2.0 * 1.0, 0.02])
    w = tf.constant([1., 0., 2., 0.])              # SUM_BY_NONZERO_WEIGHTS: sum(loss * w) / count(w != 0)
    assert np.isclose(float(n(tf.losses.huber_loss(lab, pred, weights=w))), (0.125 * 1 + 2.5 * 2) / 2)
    assert np.isclose(float(n(tf.losses.mean_squared_error(lab, pred))), np.mean([0.25, 1.0, 9.0, 0.04]))


def test_cross_entropies(tf):
    logits = np.array([[2.0, 1.0, 0.1], [0.0, 0.0, 0.0]])
    p = np.exp(logits) / np.exp(logits).sum(1, keepdims=True)
    got = n(tf.nn.sparse_softmax_cross_entropy_with_logits(logits=tf.constant(logits), labels=tf.constant([0, 2])))
    assert np.allclose(got, [-np.log(p[0, 0]), np.log(3.0)])
    soft = n(tf.nn.softmax_cross_entropy_with_logits(labels=tf.constant([[0.5, 0.5, 0.0], [0., 0., 1.]]), logits=tf.constant(logits)))
    assert np.allclose(soft, [-(0.5 * np.log(p[0, 0]) + 0.5 * np.log(p[0, 1])), np.log(3.0)])
    x, z = np.array([-3.0, 0.0, 2.5]), np.array([0.0, 1.0, 1.0])
    sig = n(tf.nn.sigmoid_cross_entropy_with_logits(labels=tf.constant(z), logits=tf.constant(x)))
    s = 1 / (1 + np.exp(-x))
    assert np.allclose(sig, -(z * np.log(s) + (1 - z) * np.log(1 - s)))


def test_batch_norm_training_and_inference(tf):
    rng = np.random.RandomState(0)
    x = rng.randn(6, 4, 1, 3)
    vals = {'s/bn/beta': np.array([0.1, -0.2, 0.3]), 's/bn/gamma': np.array([1.5, 0.5, 2.0]), 's/bn/moving_mean': np.array([1., 2., 3.]),
            's/bn/moving_variance': np.array([4., 5., 6.])}
    tf1_shim.STATE.reset({k: v.copy() for k, v in vals.items()})
    with tf.variable_scope('s'):
        y_eval = n(tf.contrib.layers.batch_norm(tf.constant(x), center=True, scale=True, is_training=tf.constant(False), decay=0.9,
                                                updates_collections=None, scope='bn'))
    want = (x - vals['s/bn/moving_mean']) / np.sqrt(vals['s/bn/moving_variance'] + 1e-3) * vals['s/bn/gamma'] + vals['s/bn/beta']
    assert np.allclose(y_eval, want)
    tf1_shim.STATE.reset({k: v.copy() for k, v in vals.items()})
    with tf.variable_scope('s'):
        y_tr = n(tf.contrib.layers.batch_norm(tf.constant(x), center=True, scale=True, is_training=tf.constant(True), decay=0.9,
                                              updates_collections=None, scope='bn'))
    m, v = x.reshape(-1, 3).mean(0), x.reshape(-1, 3).var(0)                       # biased variance normalises the batch
    assert np.allclose(y_tr, (x - m) / np.sqrt(v + 1e-3) * vals['s/bn/gamma'] + vals['s/bn/beta'])
    cnt = 24
    assert np.allclose(n(tf1_shim.STATE.vars['s/bn/moving_mean']), 0.9 * vals['s/bn/moving_mean'] + 0.1 * m)
    assert np.allclose(n(tf1_shim.STATE.vars['s/bn/moving_variance']), 0.9 * vals['s/bn/moving_variance'] + 0.1 * v * cnt / (cnt - 1))


def test_conv2d_one_by_d_kernel_is_a_per_point_linear_map(tf):
    rng = np.random.RandomState(1)
    x, w = rng.randn(2, 5, 6, 1), rng.randn(1, 6, 1, 4)          # the (B, N, D, 1) image with a [1, D] kernel: tf_util.conv2d, first layer
    y = n(tf.nn.conv2d(tf.constant(x), tf.constant(w), [1, 1, 1, 1], padding='VALID'))
    assert y.shape == (2, 5, 1, 4) and np.allclose(y[:, :, 0, :], x[:, :, :, 0] @ w[0, :, 0, :])
    x1, w1 = rng.randn(2, 5, 1, 3), rng.randn(1, 1, 3, 4)
    y1 = n(tf.nn.conv2d(tf.constant(x1), tf.constant(w1), [1, 1, 1, 1], padding='VALID'))
    assert np.allclose(y1[:, :, 0, :], x1[:, :, 0, :] @ w1[0, 0])
    pooled = n(tf.nn.max_pool(tf.constant(x1), ksize=[1, 5, 1, 1], strides=[1, 2, 2, 1], padding='VALID'))
    assert pooled.shape == (2, 1, 1, 3) and np.allclose(pooled[:, 0, 0, :], x1[:, :, 0, :].max(1))


def test_shape_protocol_slices_and_integer_division(tf):
    x = tf.constant(np.arange(24.).reshape(2, 3, 4))
    assert x.get_shape().as_list() == [2, 3, 4] and x.get_shape()[1].value == 3 and x.shape[2].value == 4
    assert n(tf.slice(x, [0, 1, 0], [-1, 1, 2])).tolist() == [[[4., 5.]], [[16., 17.]]]
    assert n(x[:, 1, 0:2]).tolist() == [[4., 5.], [16., 17.]]
    assert n(tf.tile(tf.expand_dims(tf.constant([1., 2.]), 0), [2, 2])).tolist() == [[1., 2., 1., 2.]] * 2
    assert n(tf.constant([7, -7]) / tf.constant([2, 2])).tolist() == [3, -4]          # TF1 `/` on int32 = floor division
    assert n(tf.cast(tf.constant([1.9, -1.9]), tf.int32)).tolist() == [1, -1]          # float -> int truncates
    assert n(tf.reduce_sum(x, axis=[1, 2], keep_dims=True)).shape == (2, 1, 1)
    assert np.isclose(float(n(tf.norm(tf.constant([[3., 4.]]), axis=-1))[0]), 5.0)
    assert n(tf.argmax(tf.constant([[1., 9., 9.]]), axis=1)).tolist() == [1]            # first maximum


def test_map_fn_nested_structures_and_gather_nd(tf):
    a, b = tf.constant(np.arange(6.).reshape(3, 2)), tf.constant([10., 20., 30.])
    s, d = tf.map_fn(lambda e: [e[0][0] + e[1], e[0][1] - e[1]], [(a, a * 2.0), b], dtype=[tf.float32, tf.float32])
    assert n(s).tolist() == [[10., 11.], [22., 23.], [34., 35.]] and n(d).tolist() == [[-10., -8.], [-16., -14.], [-22., -20.]]
    params = tf.constant(np.arange(12.).reshape(2, 3, 2))
    idx = tf.constant(np.array([[[0, 2], [0, 0]], [[1, 1], [1, 2]]]), dtype=tf.int32)
    assert n(tf.gather_nd(params, idx)).tolist() == [[[4., 5.], [0., 1.]], [[8., 9.], [10., 11.]]]
    assert n(tf.gather(tf.constant([5., 6., 7.]), [2, 0])).tolist() == [7., 5.]


def test_variable_scopes_collections_and_float32_constants(tf):
    tf1_shim.STATE.reset({'a/b/weights': np.ones((2, 3)), 'a/b/bn/moving_mean': np.zeros(3), 'c/weights': np.ones(1)})
    with tf.variable_scope('a'):
        with tf.variable_scope('b') as sc:
            w = tf.get_variable('weights', [2, 3])
            assert sc.name == 'a/b' and w.name == 'a/b/weights:0' and w.op.name == 'a/b/weights'
            with tf.variable_scope('bn'):
                tf.get_variable('moving_mean', [3], trainable=False)
    with tf.variable_scope('c'):
        tf.get_variable('weights', [1])
    names = [v.name for v in tf.get_collection(tf.GraphKeys.TRAINABLE_VARIABLES, scope='a')]
    assert names == ['a/b/weights:0']                                   # prefix match; the moving statistic is not trainable
    assert len(tf.get_collection(tf.GraphKeys.GLOBAL_VARIABLES)) == 3
    with pytest.raises(KeyError):
        tf.get_variable('missing', [1])
    c = n(tf.constant(np.array([0.1, 2.114256]), dtype=tf.float32))         # a tf.float32 constant holds float32 values
    assert c.dtype == np.float64 and c.tolist() == np.array([0.1, 2.114256], dtype=np.float32).astype(np.float64).tolist()
