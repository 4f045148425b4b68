"""Flag object `c` / FLAGS threaded through the model and loss functions.

Mirrors the module constants of models/config.py:13-186 (same UPPER_CASE names and defaults);
only the attributes the hot path reads are kept.  `cfg(**overrides)` returns a namespace
that plays the role of the parsed FLAGS object (models/config.py:243-350).
"""
from types import SimpleNamespace
import numpy as np

_DEFAULTS = dict(
    BOX_PC_MASK_REPRESENTATION='A',
    USE_NORMALIZED_BOX2D_AS_FEATS=False,
    NORMALIZE_PC_BEFORE_SEG=False,
    NORMALIZATION_METHOD='',
    BOXPC_NOFIT_BOUNDS=[0.01, 0.25],
    BOXPC_FIT_BOUNDS=[0.7, 1.0],
    BOXPC_CENTER_PERTURBATION=0.8,
    BOXPC_SIZE_PERTURBATION=0.2,
    BOXPC_ANGLE_PERTURBATION=np.pi,
    BOXPC_DELTA_LOSS_TYPE='huber',
    BOXPC_WEIGH_DELTA_PRED_BY_CLS_CONF=False,
    BOXPC_WEIGH_DELTA_LOSS_BY_CLS_CONF=False,
    BOXPC_STOP_GRAD_OF_CLS_VIA_DELTA=True,
    BOXPC_WEIGH_DELTA_LOSS_BY_CLS_GT=False,
    BOXPC_WEIGHT_CLS=1.,
    BOXPC_WEIGHT_DELTA=1.,
    BOXPC_WEIGHT_DELTA_CENTER_PERCENT=0.34,
    BOXPC_WEIGHT_DELTA_SIZE_PERCENT=0.33,
    BOXPC_WEIGHT_DELTA_ANGLE_PERCENT=0.33,
    SEMI_MODEL='F',
    SEMI_ADV_DROPOUTS_FOR_G=0.5,
    SEMI_ADV_TANH_FOR_LAST_LAYER_OF_G=True,
    SEMI_ADV_LEAKY_RELU=True,
    SEMI_TRAIN_BOXPC_MODEL=False,
    SEMI_TRAIN_BOX_TRAIN_CLASS_AG_TNET=False,
    SEMI_TRAIN_BOX_TRAIN_CLASS_AG_BOX=False,
    SEMI_INTRACLSDIMS_ONLY_ON_2D_CLS=True,
    WEAK_INACTIVE_VOL_ONLY_ON_2D_CLS=True,
    TEST_CLS=['table', 'sofa', 'dresser', 'night_stand', 'bookshelf'],     # SUNRGBD_SEMI_TEST_CLS (config.py:193-194)
    use_one_hot_boxpc=False,
    SEMI_BOXPC_MIN_FIT_LOSS_AFT_REFINE=False,
    SEMI_BOXPC_FIT_ONLY_ON_2D_CLS=False,
    SEMI_WEIGH_BOXPC_DELTA_DURING_TEST=False,
    SEMI_REFINE_USING_BOXPC_DELTA_NUM=1,
    SEMI_MULTIPLIER_FOR_WEAK_LOSS=1,
    SEMI_WEIGHT_BOXPC_FIT_LOSS=1.,
    WEAK_WEIGHT_INACTIVE_VOLUME=0,
    WEAK_WEIGHT_REPROJECTION=0.01,
    WEAK_WEIGHT_SURFACE=1.,
    WEAK_TRAIN_SEG_W_SURFACE=False,
    WEAK_TRAIN_BOX_W_SURFACE=[True, False, True],
    WEAK_SURFACE_MARGIN=0,
    WEAK_SURFACE_LOSS_WT_FOR_INNER_PTS=0.8,
    WEAK_SURFACE_LOSS_SCALE_DIMS=0.9,
    WEAK_WEIGHT_INTRACLASSVAR=0,
    WEAK_TRAIN_BOX_W_REPROJECTION=[True, True, True],
    WEAK_REPROJECTION_USE_SOFTMAX_PROJ=False,
    WEAK_REPROJECTION_SOFTMAX_SCALE=10.,
    WEAK_REPROJECTION_ONLY_ON_2D_CLS=False,
    WEAK_REPROJECTION_CLIP_LOWERB_LOSS=True,
    WEAK_REPROJECTION_CLIP_PRED_BOX=False,
    WEAK_REPROJECTION_LOSS_TYPE='huber',
    WEAK_REPROJECTION_DILATE_FACTOR=1.5,
    WEAK_DIMS_LOSS_TYPE='huber',
    WEAK_DIMS_USE_MARGIN_LOSS=True,
    WEAK_DIMS_SD_MARGIN=0.2,
    WEAK_INACTIVE_VOL_LOSS_MARGINS=[10., 0., 0.],
    STRONG_WEIGHT_CROSS_ENTROPY=1.,
    STRONG_BOX_MULTIPLER=0.1,
    STRONG_WEIGHT_CENTER=1.,
    STRONG_WEIGHT_ORIENT_CLS=1.,
    STRONG_WEIGHT_ORIENT_REG=20.,
    STRONG_WEIGHT_DIMS_CLS=1.,
    STRONG_WEIGHT_DIMS_REG=20.,
    STRONG_WEIGHT_TNET_CENTER=1.,
    STRONG_WEIGHT_CORNER=1.,
    # driver-level flags (test_semisup.py / train_*.py argparse, lower case there)
    use_one_hot=True,
    refine=1,
    mask_pc_for_boxpc=False,
)


def cfg(**overrides):
    d = dict(_DEFAULTS)
    for k in overrides:
        if k not in d:
            raise AttributeError('unknown flag %s' % k)
    d.update(overrides)
    return SimpleNamespace(**d)
