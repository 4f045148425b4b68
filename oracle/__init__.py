"""CPU oracle: a restatement of the reference's TF1 graph for the Frustum-PointNet hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  The product package
(transferable3d_b200) never imports it and has no CPU fallback.

PARITY UNPINNED: the reference (yewsiang/Transferable3D) ships no tests, golden vectors,
fixtures or checkpoints, and cannot be imported here (Python-2 + TensorFlow<=1.15 with
tf.contrib; neither is installable offline; its own `box_util` module is missing from the
tree).  The arithmetic lives in TensorFlow 1.x (version unpinned by the reference).  Every
function below cites the reference file:line it restates; the only numeric example the
reference holds (softmax weight table, models/config.py:137-142) is checked in
tests/test_oracle_geometry.py.

Engine: PyTorch-CPU, float32 by default (float64 switch to measure the oracle's own noise
floor), autograd available for backward parity.  Per-point tensors are kept as (B,N,C); the
reference's NHWC (B,N,1,C) singleton axis is dropped.  Anything random in the reference
(dropout masks, resampling RNG) is an explicit input here.
"""
from . import tf_layers, tf_util, model_util, semisup_models, semisup_v1_sunrgbd  # noqa: F401
from . import boxpc_sunrgbd, weak_losses, test_semisup  # noqa: F401
