"""-m "not gpu": the C-ABI library loads and exports every symbol include/t3d_b200.h declares
(no compute calls), argument validation returns negative codes without touching a device, and
the product path fails loudly without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 't3d_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(t3d_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported_and_bound(built_lib):
    from transferable3d_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), 'symbol %s declared in the header is not exported' % n
        assert n in _lib.SIGNATURES, 'symbol %s has no ctypes signature' % n
    assert set(_lib.SIGNATURES) == set(names)
    assert built_lib.t3d_version() == 100


def test_sizes_and_error_codes_without_device(built_lib):
    lib = built_lib
    assert lib.t3d_chain_arena_bytes(0) == 19 * 16384 + 4 * (6 * 64 + 64 + 256 + 1024)
    assert lib.t3d_chain_arena_bytes(1) == 6 * 16384 + 4 * (3 * 128 + 128 + 128 + 256)
    assert lib.t3d_chain_arena_bytes(2) == 22 * 16384 + 4 * (3 * 128 + 128 + 384 + 512)
    assert lib.t3d_chain_arena_bytes(3) == 22 * 16384 + 4 * (12 * 128 + 128 + 384 + 512)
    assert lib.t3d_chain_arena_bytes(9) == 0
    assert [lib.t3d_chain_tile_points(k) for k in range(4)] == [256, 256, 128, 128]
    assert [lib.t3d_chain_out_channels(k) for k in range(4)] == [1024, 256, 512, 512]
    assert lib.t3d_seg2_arena_bytes() == 26 * 16384 + 4 * 770
    # null pointers / bad shapes are rejected before any CUDA call
    assert lib.t3d_linear_f32(None, 0, None, 0, None, None, 0, None, 0, 1, 1, 1, 0, None, None, None) == -1
    assert lib.t3d_mask_centroid(None, None, 1, 1, 3, None, None, None, None, None, None) == -1
    assert lib.t3d_chain_max_bf16(7, None, 1, 1, 6, None, None, 0, None, None, None, None, None, None, None, None, None, None) == -1
    assert b'invalid argument' in lib.t3d_error_string(-1)


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_product_path_fails_loudly_without_gpu(built_lib):
    from transferable3d_b200 import runtime as rt, _lib
    with pytest.raises(_lib.T3DError):
        rt.linear(torch.zeros(4, 4), torch.zeros(4, 4))


def test_zipped_pickle_roundtrip_protocol2(tmp_path):
    """utils.save_zipped_pickle / load_zipped_pickle (sunrgbd_data/utils.py:341-348): protocol-2 gz pickle (what Python 2's
    cPickle wrote), numpy arrays and str survive the latin1 load."""
    import numpy as np
    from transferable3d_b200 import utils
    obj = [[1, 2], [np.arange(6.0).reshape(2, 3), np.ones((0, 3))], ['bed', 'night_stand']]
    p = os.path.join(str(tmp_path), 'x.zip.pickle')
    utils.save_zipped_pickle(obj, p)
    back = utils.load_zipped_pickle(p)
    assert back[0] == [1, 2] and back[2] == ['bed', 'night_stand']
    assert np.array_equal(back[1][0], obj[1][0]) and back[1][1].shape == (0, 3)
