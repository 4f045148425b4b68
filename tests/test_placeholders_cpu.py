"""-m "not gpu": the placeholder tuples of the three reference entry points -- order, shapes and
dtypes exactly as the reference builds them (semisup_v1_sunrgbd.py:37-67, boxpc_sunrgbd.py:33-54, models/model_util.py:
216-238) -- and the positional signature of get_strong_loss (:423).  Allocation only, no compute."""
import inspect

import torch

from transferable3d_b200 import semisup_v1_sunrgbd as M, boxpc_sunrgbd as BP, model_util as MU

F, I = torch.float32, torch.int32


def spec(ts):
    return [(tuple(t.shape), t.dtype) for t in ts]


def test_semisup_placeholder_inputs_a1():
    B, N, C = 5, 64, 6
    got = spec(M.placeholder_inputs(B, N, C, device='cpu'))
    # pc, bg_pc, img, one_hot | labels, centers, y_orient_cls, y_orient_reg, y_dims_cls, y_dims_reg |
    # R0_rect, P, Rtilt, K | rot_frust, box2D, img_dim, is_data_2D
    want = [((B, N, C), F), ((B, N, C), F), ((B, 0, 0, 3), F), ((B, 10), F),
            ((B, N), I), ((B, 3), F), ((B,), I), ((B,), F), ((B,), I), ((B, 3), F),
            ((B, 3, 3), F), ((B, 3, 4), F), ((B, 3, 3), F), ((B, 3, 3), F),
            ((B, 1), F), ((B, 4), F), ((B, 2), F), ((B,), I)]
    assert got == want and len(got) == 18


def test_boxpc_placeholder_inputs_a2():
    B, N, C = 3, 32, 6
    got = spec(BP.placeholder_inputs(B, N, C, device='cpu'))
    # pc, one_hot, y_seg, x_center, x_orient_cls, x_orient_reg, x_dims_cls, x_dims_reg, y_box_iou, y_center_delta,
    # y_dims_delta, y_orient_delta  -- dims delta BEFORE orient delta (boxpc_sunrgbd.py:52-54)
    want = [((B, N, C), F), ((B, 10), F), ((B, N), I), ((B, 3), F), ((B,), I), ((B,), F), ((B,), I), ((B, 3), F),
            ((B,), F), ((B, 3), F), ((B, 3), F), ((B,), F)]
    assert got == want and len(got) == 12


def test_model_util_placeholder_inputs_a3():
    B, N = 4, 128
    got = spec(MU.placeholder_inputs(B, N, device='cpu'))
    # KITTI widths of the inherited helper: 4-channel points, 3-class one-hot
    want = [((B, N, 4), F), ((B, 3), F), ((B, N), I), ((B, 3), F), ((B,), I), ((B,), F), ((B,), I), ((B, 3), F)]
    assert got == want
    got = spec(MU.placeholder_inputs(B, N, num_channel=6, num_class=10, device='cpu'))
    assert got[0] == ((B, N, 6), F) and got[1] == ((B, 10), F)


def test_get_strong_loss_signature_matches_reference():
    """reference: get_strong_loss(pred, labels, end_points, prefix='', reg_weight=0.001, reduce_loss=True, c=None)"""
    from oracle import semisup_v1_sunrgbd as OM
    for fn in (M.get_strong_loss, OM.get_strong_loss):
        p = inspect.signature(fn).parameters
        assert list(p) == ['pred', 'labels', 'end_points', 'prefix', 'reg_weight', 'reduce_loss', 'c']
        assert p['prefix'].default == '' and p['reg_weight'].default == 0.001 and p['reduce_loss'].default is True
