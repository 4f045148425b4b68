"""transferable3d_b200: B200-native (sm_100a) Frustum-PointNet hot path of yewsiang/Transferable3D.

Host mirror of the reference's model entry points (same names / arguments / end_points keys):
  semisup_models, semisup_v1_sunrgbd, boxpc_sunrgbd, model_util, tf_util, test_semisup,
  frustum_pointnets_v1 (cfg3 pipeline), config (FLAGS), weights / synth (synthetic data).
All compute runs in hand-written CUDA kernels behind the C ABI of include/t3d_b200.h
(libt3d_b200.so); there is no CPU fallback.
"""
from . import constants, config, weights, synth  # noqa: F401  (pure numpy; importable without a GPU)


def _lazy(name):
    import importlib
    return importlib.import_module('.' + name, __name__)


def __getattr__(name):
    if name in ('runtime', 'tf_util', 'semisup_models', 'semisup_v1_sunrgbd', 'boxpc_sunrgbd', 'model_util',
                'test_semisup', 'frustum_pointnets_v1', '_lib'):
        return _lazy(name)
    raise AttributeError(name)
