"""Host-side runtime of the B200 path: the variable store (stands in for the TF variable
collection + tf.variable_scope, models/tf_util.py:1151-1190), BN folding / weight packing caches,
the precision switch, and thin wrappers over the C ABI.

Precision modes (BASELINE north_star):
  'bf16'   fused tcgen05 kernels, bf16 operands / fp32 accumulate (the throughput mode);
  'f16x2'  the same fused kernels in split precision: every operand is an fp16 hi + lo pair, three tensor-core products
           per layer, fp32-accurate (the mode that meets the mask-exactness target; csrc/chain_x2.cuh);
  'fp32'   layer-by-layer fp32 GEMMs with the literal mask multiply (1e-4 parity mode, every activation through HBM).
FC heads are fp32 in all three.  Eval-mode graphs only; training graphs live in train_boxpc / train_semisup_adv.
"""
import contextlib
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import ptr, stream, call
from .weights import fold_bn

ACT = {'none': 0, None: 0, 'relu': 1, 'leaky_relu': 2, 'tanh': 3}
CHAIN_SEG1, CHAIN_TNET, CHAIN_BOX, CHAIN_BOXPC, CHAIN_BOXPCB = 0, 1, 2, 3, 4

PRECISIONS = ('bf16', 'f16x2', 'fp32')
FUSED = ('bf16', 'f16x2')          # modes that run the fused tcgen05 chains
X2_ACT_SCALE = 16.0               # csrc/chain_x2.cuh: kX2ActScale

_state = {'store': None, 'precision': 'bf16'}


def set_precision(mode):
    if mode not in PRECISIONS:
        raise ValueError(mode)
    _state['precision'] = mode


def get_precision():
    return _state['precision']


def is_x2():
    return _state['precision'] == 'f16x2'


def set_x2_debias(per_step):
    """Correction of the tensor core's truncating accumulation folded into the f16x2 epilogue scales (include/t3d_b200.h:
    t3d_set_x2_debias).  Arenas are cached per value."""
    _lib.check(_lib.load().t3d_set_x2_debias(float(per_step)))


def get_x2_debias():
    return float(_lib.load().t3d_get_x2_debias())


def _pow2_scale(w):
    """Power of two s with max|w| * s in [2^9, 2^10): keeps the fp16 lo image of the weights in the normal range."""
    m = float(w.abs().max())
    if not np.isfinite(m) or m <= 0.0:
        return 1.0
    return float(2.0 ** (9 - int(np.floor(np.log2(m)))))


def default_device():
    """Device of the default variable store, else the current CUDA device (raises without one: no CPU fallback)."""
    if _state['store'] is not None:
        return _state['store'].device
    if not torch.cuda.is_available():
        raise _lib.T3DError('t3d_b200 ops need a CUDA device (no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def set_f32_engine(engine):
    """Engine of the fp32 GEMMs (fp32-mode layers, training steps): 'tc' = tcgen05 bf16 x 3 split (fp32-accurate,
    default), 'tc2' = tcgen05 bf16 x 2 split (three products per MAC, operand error <= 2^-18: between TF32 and fp32),
    'simt' = CUDA-core SGEMM, 'bf16' = tcgen05 with operands rounded to bf16 (one pass, fp32 accumulate).
    See include/t3d_b200.h: t3d_set_f32_engine."""
    if engine not in F32_ENGINES:
        raise ValueError(engine)
    _lib.check(_lib.load().t3d_set_f32_engine(F32_ENGINES[engine]))


F32_ENGINES = {'simt': 0, 'tc': 1, 'bf16': 2, 'tc2': 3}


def get_f32_engine():
    code = _lib.load().t3d_get_f32_engine()
    return [k for k, v in F32_ENGINES.items() if v == code][0]


@contextlib.contextmanager
def f32_engine(engine):
    old = get_f32_engine()
    set_f32_engine(engine)
    try:
        yield
    finally:
        set_f32_engine(old)


@contextlib.contextmanager
def precision(mode):
    old = _state['precision']
    set_precision(mode)
    try:
        yield
    finally:
        _state['precision'] = old


class VariableStore(object):
    """Variables by TF checkpoint name (numpy, host) + device-side caches derived from them."""

    def __init__(self, variables, device=None):
        self.variables = {k: np.asarray(v) for k, v in variables.items()}
        self.device = torch.device(device if device is not None else 'cuda')
        self._scope = []
        self._folded = {}      # layer -> (W [K,N] f32 dev, b [N] f32 dev)
        self._arenas = {}      # (key) -> packed uint8 dev tensor
        self._consts = {}

    # --- scopes ---------------------------------------------------------------------------
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

    # --- derived device tensors -------------------------------------------------------------
    def folded(self, layer):
        """(W[Cin,Cout], b[Cout]) fp32 on device with eval-mode BN folded in."""
        if layer not in self._folded:
            if layer + '/weights' not in self.variables:
                raise KeyError('variable %s/weights not found' % layer)
            w, b = fold_bn(self.variables, layer)
            self._folded[layer] = (torch.from_numpy(w).to(self.device), torch.from_numpy(b).to(self.device))
        return self._folded[layer]

    def const(self, name, array):
        if name not in self._consts:
            self._consts[name] = torch.as_tensor(np.asarray(array, dtype=np.float32)).to(self.device).contiguous()
        return self._consts[name]

    def cached(self, key, make):
        """Device-side value derived from the variables, built once (`make()` returns tensors on self.device)."""
        if key not in self._consts:
            self._consts[key] = make()
        return self._consts[key]

    def _new_arena(self, nbytes):
        # 1024-byte aligned base (torch allocations are 512-byte aligned)
        raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
        off = (-raw.data_ptr()) % 1024
        return raw[off:off + nbytes]

    def chain_arena(self, scope, kind, layers, x2=False):
        """Packed tcgen05 weights of one per-point chain. `layers` = TF layer names under `scope`:
        [layer1, hidden..., final].  x2: fp16 hi / lo images for the f16x2 kernels."""
        key = ('chain', scope, kind, get_x2_debias() if x2 else None)
        if key not in self._arenas:
            lib = _lib.load()
            assert lib.t3d_chain_num_layers(kind) == len(layers)
            ws, bs = [], []
            for name in layers:
                w, b = self.folded('%s/%s' % (scope, name))
                ws.append(w.contiguous())
                bs.append(b.contiguous())
            PA = ctypes.c_void_p * len(layers)
            if x2:
                arena = self._new_arena(lib.t3d_chain_arena_bytes_x2(kind))
                scales = (ctypes.c_float * (len(layers) - 1))(*[_pow2_scale(w) for w in ws[1:]])
                call('t3d_pack_chain_x2', kind, PA(*[w.data_ptr() for w in ws]), PA(*[b.data_ptr() for b in bs]), scales,
                     ptr(arena), stream())
            else:
                arena = self._new_arena(lib.t3d_chain_arena_bytes(kind))
                call('t3d_pack_chain', kind, PA(*[w.data_ptr() for w in ws]), PA(*[b.data_ptr() for b in bs]),
                     ptr(arena), stream())
            torch.cuda.current_stream().synchronize()       # ws/bs may be freed after this
            self._arenas[key] = arena
        return self._arenas[key]

    def seg2_arena(self, scope, x2=False):
        if x2:
            key = ('seg2x', scope, get_x2_debias())
            if key not in self._arenas:
                lib = _lib.load()
                w6, _ = self.folded(scope + '/conv6')
                w6p = w6[:64].contiguous()
                (w7, b7), (w8, b8), (w9, b9) = [self.folded(scope + '/conv%d' % i) for i in (7, 8, 9)]
                w10, b10 = self.folded(scope + '/conv10')
                scales = (ctypes.c_float * 4)(*[_pow2_scale(w) for w in (w6p, w7, w8, w9)])
                arena = self._new_arena(lib.t3d_seg2_arena_bytes_x2())
                call('t3d_pack_seg2_x2', ptr(w6p), ptr(w7), ptr(w8), ptr(w9), ptr(b7), ptr(b8), ptr(b9), ptr(w10), ptr(b10),
                     scales, ptr(arena), stream())
                torch.cuda.current_stream().synchronize()
                self._arenas[key] = arena
            return self._arenas[key]
        key = ('seg2', scope)
        if key not in self._arenas:
            lib = _lib.load()
            w6, _ = self.folded(scope + '/conv6')
            w6p = w6[:64].contiguous()
            w7, b7 = self.folded(scope + '/conv7')
            w8, b8 = self.folded(scope + '/conv8')
            w9, b9 = self.folded(scope + '/conv9')
            w10, b10 = self.folded(scope + '/conv10')
            arena = self._new_arena(lib.t3d_seg2_arena_bytes())
            call('t3d_pack_seg2', ptr(w6p), ptr(w7), ptr(w8), ptr(w9), ptr(b7), ptr(b8), ptr(b9), ptr(w10), ptr(b10),
                 ptr(arena), stream())
            torch.cuda.current_stream().synchronize()
            self._arenas[key] = arena
        return self._arenas[key]


def set_default_store(store):
    _state['store'] = store


def store():
    if _state['store'] is None:
        raise RuntimeError('no VariableStore: call runtime.set_default_store(VariableStore(variables))')
    return _state['store']


@contextlib.contextmanager
def variable_scope(name):
    """tf.variable_scope analogue on the default store."""
    with store().variable_scope(name):
        yield


def require_eval(is_training):
    if bool(is_training):
        raise NotImplementedError('training-mode (batch-statistics BN, dropout) graphs are not implemented on the '
                                  'B200 path yet; only is_training=False')


class WirePoints(object):
    """The 6-channel frustum input in its wire format (model_util.assemble_point_cloud(..., lazy=True)): xyz (B,N,3) fp32 and
    rgb (B,N,3) uint8, colours = k / 255.  The fused bf16 inst_seg chain reads it directly (t3d_chain_max_bf16_wire); every
    other consumer of the point cloud downstream of the segmentation net uses xyz only (model_util.py:241-286), and
    `dense()` builds the (B,N,6) fp32 placeholder tensor for the paths that need it (f16x2 / fp32 precision modes)."""

    def __init__(self, xyz, rgb):
        self.xyz, self.rgb = xyz, rgb
        self.shape = (xyz.shape[0], xyz.shape[1], 6)
        self.device = xyz.device
        self._dense = None

    def dense(self):
        if self._dense is None:
            out = torch.empty(self.shape, dtype=torch.float32, device=self.device)
            call('t3d_assemble_points', ptr(self.xyz), ptr(self.rgb), self.shape[0] * self.shape[1], ptr(out), stream())
            self._dense = out
        return self._dense


def f32(t):
    if isinstance(t, WirePoints):
        return t
    return t.to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------------ op wrappers

def linear(x, w, b=None, act=None, gbias=None, rows_per_group=0, rowmask=None, gmax_groups=0, want_y=True):
    """t3d_linear_f32. x: (M,K) f32 (row stride = K), w: (K,N). Returns (y or None, gmax or None)."""
    x = f32(x)
    M, K = x.shape
    N = w.shape[1]
    assert w.shape[0] == K, (tuple(w.shape), tuple(x.shape))
    y = torch.empty((M, N), dtype=torch.float32, device=x.device) if want_y else None
    gmax = torch.zeros((gmax_groups, N), dtype=torch.float32, device=x.device) if gmax_groups else None
    call('t3d_linear_f32', ptr(x), K, ptr(w), N, ptr(b), ptr(gbias), int(rows_per_group), ptr(y), N, M, K, N,
         ACT[act], ptr(rowmask), ptr(gmax), stream())
    return y, gmax


def mask_centroid(logits, pc, want_mask=True, want_xyz_stage1=False, want_idx=True):
    B, N, C = pc.shape
    dev = pc.device
    mask = torch.empty((B, N), dtype=torch.float32, device=dev) if want_mask else None
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    mean = torch.empty((B, 3), dtype=torch.float32, device=dev)
    xyz1 = torch.empty((B, N, 3), dtype=torch.float32, device=dev) if want_xyz_stage1 else None
    idx = torch.empty((B, N), dtype=torch.int32, device=dev) if want_idx else None
    logits = f32(logits)                                          # keep alive until the launch
    call('t3d_mask_centroid', ptr(logits), ptr(pc), B, N, C, ptr(mask), ptr(count), ptr(mean), ptr(xyz1), ptr(idx),
         stream())
    return mask, count, mean, xyz1, idx


def build_tiles(count, tile_pts, max_per_frustum):
    B = count.shape[0]
    tiles = torch.empty((B * ((max_per_frustum + tile_pts - 1) // tile_pts), 4), dtype=torch.int32, device=count.device)
    num = torch.empty((1,), dtype=torch.int32, device=count.device)
    call('t3d_build_tiles', ptr(count), B, tile_pts, ptr(tiles), ptr(num), stream())
    return tiles, num


def chain_max(kind, pc, arena, center=None, idx=None, count=None, box=None, emit=None, x2=False):
    """Fused per-point chain + max on tcgen05. Returns (B, FC) fp32.  x2: the f16x2 kernel (arena packed with x2=True)."""
    lib = _lib.load()
    wire = isinstance(pc, WirePoints)
    if wire and x2:
        pc, wire = pc.dense(), False
    B, N, C = pc.shape
    fc = lib.t3d_chain_out_channels(kind)
    out = torch.empty((B, fc), dtype=torch.float32, device=pc.device)
    tiles = num = None
    idx_stride = 0
    if idx is not None:
        idx_stride = idx.shape[1]
        if count is None:
            count = torch.full((B,), idx_stride, dtype=torch.int32, device=pc.device)
        tiles, num = build_tiles(count, 128 if x2 else lib.t3d_chain_tile_points(kind), idx_stride)
    bc = bd = bo = None
    if box is not None:
        bc, bd, bo = [f32(t) for t in box]
    if wire:
        call('t3d_chain_max_bf16_wire', kind, ptr(pc.xyz), ptr(pc.rgb), B, N, ptr(center), ptr(idx), idx_stride, ptr(count), ptr(tiles), ptr(num),
             ptr(bc), ptr(bd), ptr(bo), ptr(arena), ptr(out), ptr(emit), stream())
        return out
    call('t3d_chain_max_x2' if x2 else 't3d_chain_max_bf16', kind, ptr(pc), B, N, C, ptr(center), ptr(idx), idx_stride, ptr(count), ptr(tiles), ptr(num),
         ptr(bc), ptr(bd), ptr(bo), ptr(arena), ptr(out), ptr(emit), stream())
    return out


def seg_stage2(point_feat, gbias, arena, B, N, x2=False):
    logits = torch.empty((B, N, 2), dtype=torch.float32, device=gbias.device)
    call('t3d_seg_stage2_x2' if x2 else 't3d_seg_stage2_bf16', ptr(point_feat), ptr(gbias), ptr(arena), ptr(logits), B, N, stream())
    return logits
