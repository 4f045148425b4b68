"""CPU oracle: a restatement of the reference's TF1 graph for the Frustum-PointNet hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  The product package
(transferable3d_b200) never imports it and has no CPU fallback.

PARITY PINNED AGAINST THE REFERENCE'S SOURCE, NOT AGAINST TENSORFLOW'S KERNELS.  The reference
(yewsiang/Transferable3D) ships no tests, golden vectors, fixtures or checkpoints, and TensorFlow<=1.15
(with tf.contrib) cannot be installed here; its own `box_util` module is missing from the tree.  But
its files are valid Python 3, so they are EXECUTED, unmodified and from where they lie, on a TF1
stand-in (tests/golden/tf1_shim.py: each tf.* op on PyTorch-CPU, float64): test_semisup.get_model
called as is, the graph blocks of train_boxpc.train() / train_semisup_adv.train() /
train_semisup.train() run from the scripts' ASTs, model_util / tf_util / the numpy helpers called
directly.  tests/golden/ref_*.npz hold those outputs (make_reference_golden.py) and
tests/test_oracle_vs_reference_cpu.py holds this package to them at 1e-9: losses, end points, the
trained-variable lists, every gradient, moving statistics, schedules, config defaults, the legacy
numpy resampling / perturbation streams.  What stays unpinned: the arithmetic INSIDE each TF op
(supplied by the stand-in from TF1's documented semantics: fused batch-norm's Bessel-corrected
moving variance, tf.losses' SUM_BY_NONZERO_WEIGHTS reduction, tf.where's row select, ...), and
box_util.box3d_iou (restated from the published frustum-pointnets routine).  Every function below
cites the reference file:line it restates; the only numeric example the reference holds (softmax
weight table, models/config.py:137-142) is checked in tests/test_oracle_cpu.py.

Engine: PyTorch-CPU, float32 by default (float64 switch to measure the oracle's own noise
floor), autograd available for backward parity.  Per-point tensors are kept as (B,N,C); the
reference's NHWC (B,N,1,C) singleton axis is dropped.  Anything random in the reference
(dropout masks, resampling RNG) is an explicit input here.
"""
from . import tf_layers, tf_util, model_util, semisup_models, semisup_v1_sunrgbd  # noqa: F401
from . import boxpc_sunrgbd, weak_losses, test_semisup  # noqa: F401
