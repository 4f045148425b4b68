"""A stand-in for the `tensorflow` (1.x) module, just large enough to EXECUTE the reference's own graph-building code
(/root/reference/models/{tf_util,model_util,weak_losses}.py, sunrgbd/sunrgbd_detection/{semisup_models,semisup_v1_sunrgbd,
boxpc_sunrgbd}.py) eagerly, op by op, on PyTorch-CPU tensors.  TEST INFRASTRUCTURE ONLY: used by
tests/golden/make_reference_golden.py to produce the fixtures that pin `oracle/` against the reference source, and by
tests/test_oracle_vs_reference_cpu.py when /root/reference is present.  Nothing under transferable3d_b200/ imports it.

What this pins and what it does not.  The layer order, scopes and variable names, concatenations, masks, slices, anchors,
box conversions, loss formulas, stop_gradient placement and flag handling all come from the reference's files, unmodified,
run from where they lie.  The arithmetic of each TF op is supplied here, from TensorFlow 1.x's documented semantics
(each non-obvious one says which rule it follows); TensorFlow's kernels themselves are not available in this image.

`tf.float32` maps to the module-wide float type (set_float): the fixtures are generated in float64 so that the comparison
with the oracle's float64 run is a structural identity (<= 1e-9), not a round-off budget.
"""
import contextlib
import sys
import types

import numpy as np
import torch

_FLOAT = [torch.float64]


def set_float(dt):
    _FLOAT[0] = dt


class DType(object):
    def __init__(self, name, torch_dtype, is_float=False):
        self.name, self._t, self._is_float = name, torch_dtype, is_float

    @property
    def t(self):
        return _FLOAT[0] if self._is_float else self._t

    def __eq__(self, o):
        return isinstance(o, DType) and o.name == self.name

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return 'tf.' + self.name


float32 = DType('float32', torch.float32, True)
float16 = DType('float16', torch.float32, True)
float64 = DType('float64', torch.float64)
int32 = DType('int32', torch.int32)
int64 = DType('int64', torch.int64)
bool_ = DType('bool', torch.bool)
_BY_TORCH = {torch.int32: int32, torch.int64: int64, torch.bool: bool_, torch.float32: float32, torch.float64: float32}


class Dimension(object):
    def __init__(self, v):
        self.value = v

    def __int__(self):
        return int(self.value)

    __index__ = __int__

    def __eq__(self, o):
        return self.value == (o.value if isinstance(o, Dimension) else o)

    def __hash__(self):
        return hash(self.value)

    def __repr__(self):
        return 'Dimension(%r)' % self.value

    def _v(self, o):
        return o.value if isinstance(o, Dimension) else o

    def __mul__(self, o): return self.value * self._v(o)
    __rmul__ = __mul__
    def __add__(self, o): return self.value + self._v(o)
    __radd__ = __add__
    def __sub__(self, o): return self.value - self._v(o)
    def __rsub__(self, o): return self._v(o) - self.value
    def __floordiv__(self, o): return self.value // self._v(o)
    def __truediv__(self, o): return self.value / self._v(o)


class TensorShape(object):
    def __init__(self, dims):
        self._d = [int(d) for d in dims]

    def as_list(self):
        return list(self._d)

    @property
    def ndims(self):
        return len(self._d)

    @property
    def dims(self):
        return [Dimension(d) for d in self._d]

    def __len__(self):
        return len(self._d)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return TensorShape(self._d[i])
        return Dimension(self._d[i])

    def __iter__(self):
        return iter(Dimension(d) for d in self._d)

    def __eq__(self, o):
        return self.as_list() == (o.as_list() if isinstance(o, TensorShape) else list(o))

    def __repr__(self):
        return '(%s)' % ', '.join(str(d) for d in self._d)

    __str__ = __repr__


def _int(v):
    if isinstance(v, T):
        return int(v.t)
    if isinstance(v, Dimension):
        return int(v.value)
    return int(v)


def _ints(seq):
    if isinstance(seq, T):
        return [int(v) for v in seq.t.reshape(-1).tolist()]
    if isinstance(seq, TensorShape):
        return seq.as_list()
    if isinstance(seq, (int, np.integer, Dimension)):
        return [_int(seq)]
    return [_int(v) for v in seq]


class T(object):
    """A TF1 Tensor / Variable handle around a torch tensor."""
    __array_priority__ = 1000      # numpy operands defer to the reflected operators below

    def __init__(self, t, name=None):
        self.t = t
        self.name = name or 'shim:0'

    # ---- static shape / dtype protocol -----------------------------------------------------------------------------
    @property
    def shape(self):
        return TensorShape(self.t.shape)

    def get_shape(self):
        return TensorShape(self.t.shape)

    def set_shape(self, shape):
        pass

    @property
    def dtype(self):
        return _BY_TORCH[self.t.dtype]

    @property
    def op(self):
        return types.SimpleNamespace(name=self.name.split(':')[0])

    def eval(self, *a, **k):
        return self.t.detach().numpy()

    def __repr__(self):
        return 'T(%s, %s)' % (tuple(self.t.shape), self.t.dtype)

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        return iter(T(self.t[i]) for i in range(self.t.shape[0]))

    def __bool__(self):
        return bool(self.t)

    __nonzero__ = __bool__

    def __hash__(self):
        return id(self)

    # ---- operators -------------------------------------------------------------------------------------------------
    def _o(self, o):
        return _raw(o, like=self.t)

    def __add__(self, o): return T(self.t + self._o(o))
    def __radd__(self, o): return T(self._o(o) + self.t)
    def __sub__(self, o): return T(self.t - self._o(o))
    def __rsub__(self, o): return T(self._o(o) - self.t)
    def __mul__(self, o): return T(self.t * self._o(o))
    def __rmul__(self, o): return T(self._o(o) * self.t)

    def __truediv__(self, o):
        a, b = self.t, self._o(o)
        if not a.is_floating_point():      # TF1 python-2 `/` on integers is a floor division (tf.div)
            return T(torch.div(a, b, rounding_mode='floor'))
        return T(a / b)

    def __rtruediv__(self, o): return T(self._o(o) / self.t)
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __floordiv__(self, o): return T(torch.div(self.t, self._o(o), rounding_mode='floor'))
    def __mod__(self, o): return T(torch.remainder(self.t, self._o(o)))
    def __pow__(self, o): return T(self.t ** self._o(o))
    def __rpow__(self, o): return T(self._o(o) ** self.t)
    def __neg__(self): return T(-self.t)
    def __abs__(self): return T(self.t.abs())
    def __lt__(self, o): return T(self.t < self._o(o))
    def __le__(self, o): return T(self.t <= self._o(o))
    def __gt__(self, o): return T(self.t > self._o(o))
    def __ge__(self, o): return T(self.t >= self._o(o))
    def __and__(self, o): return T(self.t & self._o(o))
    def __or__(self, o): return T(self.t | self._o(o))
    def __invert__(self): return T(~self.t)
    def __matmul__(self, o): return T(self.t @ self._o(o))

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        out = []
        for i in idx:
            if isinstance(i, T):
                i = int(i.t) if i.t.dim() == 0 else i.t.long()
            elif isinstance(i, Dimension):
                i = int(i)
            elif isinstance(i, slice):
                i = slice(*[None if v is None else _int(v) for v in (i.start, i.stop, i.step)])
            elif i is np.newaxis:
                i = None
            out.append(i)
        return T(self.t[tuple(out)])


def _raw(o, like=None, dtype=None):
    """Anything -> torch tensor (TF's convert_to_tensor with the other operand's dtype as the hint)."""
    if isinstance(o, T):
        return o.t
    if isinstance(o, Dimension):
        o = o.value
    if isinstance(o, torch.Tensor):
        return o
    if isinstance(o, (list, tuple)) and any(isinstance(v, (T, Dimension)) for v in _flatten(o)):
        return torch.stack([_raw(v, like=like, dtype=dtype) for v in o])
    if dtype is None and like is not None:
        dtype = like.dtype
    a = np.asarray(o)
    if dtype is None:
        dtype = _FLOAT[0] if a.dtype.kind == 'f' else {'i': torch.int32, 'u': torch.int32, 'b': torch.bool}[a.dtype.kind]
    return torch.as_tensor(a).to(dtype)


def _flatten(o):
    for v in o:
        if isinstance(v, (list, tuple)):
            for w in _flatten(v):
                yield w
        else:
            yield v


def _w(x):
    return x if isinstance(x, T) else T(_raw(x))


def _dt(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, DType):
        return dtype.t
    if isinstance(dtype, torch.dtype):
        return dtype
    return {'float32': _FLOAT[0], 'int32': torch.int32, 'int64': torch.int64, 'bool': torch.bool}[np.dtype(dtype).name]


# ---- variables and scopes ----------------------------------------------------------------------------------------------
class _State(object):
    def __init__(self):
        self.reset({})

    def reset(self, variables, trainable_grad=False):
        self.values = variables            # full TF name -> numpy array (the same dict the oracle's VarStore reads)
        self.vars = {}                     # full TF name -> T (created on first get_variable)
        self.scope = []
        self.collections = {}
        self.dropout_masks = {}            # full scope name -> keep mask; consumed by tf.nn.dropout
        self.dropout_log = []
        self.created = []
        self.requested_shapes = {}         # full TF name -> the shape the graph code asked for
        self.requires_grad = trainable_grad
        self.summaries = {}
        self.feeds = []


STATE = _State()


class _Scope(object):
    def __init__(self, name, reuse=None):
        self.name, self.reuse = name, reuse
        self.original_name_scope = name + '/'

    def reuse_variables(self):
        self.reuse = True


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, reuse=None, **kw):
    if isinstance(name_or_scope, _Scope):          # re-entering a captured scope: absolute
        saved = STATE.scope
        STATE.scope = [p for p in name_or_scope.name.split('/') if p]
        try:
            yield _Scope('/'.join(STATE.scope), reuse)
        finally:
            STATE.scope = saved
        return
    name = name_or_scope if name_or_scope is not None else default_name
    if name is None:
        raise ValueError('variable_scope(None)')
    STATE.scope.append(name)
    try:
        yield _Scope('/'.join(STATE.scope), reuse)
    finally:
        STATE.scope.pop()


name_scope = variable_scope


def get_variable_scope():
    return _Scope('/'.join(STATE.scope))


def _full(name):
    return '/'.join([p for p in STATE.scope if p] + [name])


def get_variable(name, shape=None, initializer=None, dtype=None, trainable=True, **kw):
    full = _full(name)
    if full in STATE.vars:
        return STATE.vars[full]
    if full not in STATE.values:
        raise KeyError('reference asked for variable %r, which the weight set does not hold' % full)
    v = torch.as_tensor(np.asarray(STATE.values[full])).to(_dt(dtype) or _FLOAT[0]).clone()
    if shape is not None:
        shape = _ints(shape)
        STATE.requested_shapes[full] = list(shape)
        assert int(np.prod(shape)) == v.numel(), (full, shape, tuple(v.shape))
        v = v.reshape(shape)
    if STATE.requires_grad and trainable and v.is_floating_point():
        v.requires_grad_(True)
    var = T(v, name=full + ':0')
    STATE.vars[full] = var
    STATE.created.append(full)
    STATE.collections.setdefault(GraphKeys.GLOBAL_VARIABLES, []).append(var)
    if trainable:
        STATE.collections.setdefault(GraphKeys.TRAINABLE_VARIABLES, []).append(var)
    return var


def Variable(initial_value, trainable=True, name=None, **kw):
    full = _full(name or 'Variable')
    if full in STATE.values:
        return get_variable(name or 'Variable', trainable=trainable)
    var = T(_raw(initial_value).clone(), name=full + ':0')
    STATE.vars[full] = var
    STATE.collections.setdefault(GraphKeys.GLOBAL_VARIABLES, []).append(var)
    if trainable:
        STATE.collections.setdefault(GraphKeys.TRAINABLE_VARIABLES, []).append(var)
    return var


def constant_initializer(value=0.0, **kw):
    return ('constant', value)


def truncated_normal_initializer(**kw):
    return ('truncated_normal', kw)


def add_to_collection(name, value):
    STATE.collections.setdefault(name, []).append(value)


def get_collection(name, scope=None):
    """`scope` filters with re.match on the item's name, i.e. a prefix match (tf.get_collection)."""
    items = list(STATE.collections.get(name, []))
    if scope is None:
        return items
    import re
    return [v for v in items if hasattr(v, 'name') and re.match(scope, v.name)]


@contextlib.contextmanager
def _noop_ctx(*a, **k):
    yield


device = control_dependencies = _noop_ctx


def no_op(*a, **k):
    return None


def group(*a, **k):
    return None


# ---- creation / shape ops ---------------------------------------------------------------------------------------------
def placeholder(dtype, shape=None, name=None):
    """Eager stand-in: the n-th placeholder a graph-building function creates takes the n-th value of STATE.feeds (None: a
    zero tensor of the declared shape, unknown dimensions as 1)."""
    if not STATE.feeds:
        raise RuntimeError('tf.placeholder: no feed value queued (the shim runs eagerly)')
    v = STATE.feeds.pop(0)
    if v is None:
        return T(torch.zeros([1 if d is None else int(d) for d in (shape or [])], dtype=_dt(dtype)))
    t = _raw(v, dtype=_dt(dtype))
    if shape is not None:
        assert len(shape) == t.dim() and all(d is None or int(d) == s for d, s in zip(shape, t.shape)), (shape, tuple(t.shape))
    return T(t)


class Graph(object):
    def as_default(self):
        return _noop_ctx()


class _Anything(object):
    """ConfigProto / Session / Saver: attribute sinks (nothing to configure, restore or run in the eager stand-in)."""
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Anything()

    def __setattr__(self, name, value):
        pass

    def __call__(self, *a, **k):
        return _Anything()


ConfigProto = Session = _Anything


class GraphKeys(object):
    TRAINABLE_VARIABLES, GLOBAL_VARIABLES, UPDATE_OPS = 'trainable_variables', 'variables', 'update_ops'


def constant(value, dtype=None, shape=None, name=None):
    t = _raw(value, dtype=_dt(dtype))
    if isinstance(dtype, DType) and dtype._is_float and t.dtype == torch.float64:
        t = t.to(torch.float32).to(torch.float64)      # a tf.float32 constant holds float32 values, whatever the run's float type
    if shape is not None:
        t = t.expand(_ints(shape)).clone() if t.numel() == 1 else t.reshape(_ints(shape))
    return T(t)


def convert_to_tensor(value, dtype=None, name=None, **kw):
    if isinstance(value, T) and dtype is None:
        return value
    return T(_raw(value, dtype=_dt(dtype)))


def zeros(shape, dtype=float32, name=None): return T(torch.zeros(_ints(shape), dtype=_dt(dtype)))
def ones(shape, dtype=float32, name=None): return T(torch.ones(_ints(shape), dtype=_dt(dtype)))
def zeros_like(x, dtype=None, **kw): return T(torch.zeros_like(_raw(x), dtype=_dt(dtype)))
def ones_like(x, dtype=None, **kw): return T(torch.ones_like(_raw(x), dtype=_dt(dtype)))
def identity(x, name=None): return _w(x)
def stop_gradient(x, name=None): return T(_raw(x).detach())
def to_float(x, name=None): return T(_raw(x).to(_FLOAT[0]))
def to_int32(x, name=None): return T(_raw(x).to(torch.int32))
def to_int64(x, name=None): return T(_raw(x).to(torch.int64))


def cast(x, dtype, name=None):
    t, d = _raw(x), _dt(dtype)
    if t.is_floating_point() and not d.is_floating_point and d != torch.bool:
        t = torch.trunc(t)
    return T(t.to(d))


def shape(x, name=None, out_type=int32):
    return T(torch.tensor(list(_raw(x).shape), dtype=torch.int32))


def rank(x): return T(torch.tensor(_raw(x).dim(), dtype=torch.int32))
def size(x): return T(torch.tensor(_raw(x).numel(), dtype=torch.int32))


def range_(start, limit=None, delta=1, dtype=None, name=None):
    if limit is None:
        start, limit = 0, start
    vals = [start, limit, delta]
    is_f = any(isinstance(v, float) or (isinstance(v, T) and v.t.is_floating_point()) for v in vals)
    a, b, c = [(float(_raw(v)) if is_f else _int(v)) for v in vals]
    d = _dt(dtype) or (_FLOAT[0] if is_f else torch.int32)
    return T(torch.arange(a, b, c).to(d))


def reshape(x, shape, name=None):
    return T(_raw(x).reshape(_ints(shape)))


def expand_dims(x, axis=None, name=None, dim=None):
    return T(_raw(x).unsqueeze(_int(axis if axis is not None else dim)))


def squeeze(x, axis=None, name=None, squeeze_dims=None):
    t = _raw(x)
    axis = axis if axis is not None else squeeze_dims
    if axis is None:
        return T(t.squeeze())
    ax = sorted([a % t.dim() for a in _ints(axis)], reverse=True)
    for a in ax:
        assert t.shape[a] == 1, ('squeeze of a non-unit axis', tuple(t.shape), a)
        t = t.squeeze(a)
    return T(t)


def tile(x, multiples, name=None):
    t = _raw(x)
    m = _ints(multiples)
    assert len(m) == t.dim(), ('tf.tile needs one multiple per axis', tuple(t.shape), m)
    return T(t.repeat(*m))


def concat(values=None, axis=None, name=None, **kw):
    if isinstance(values, (int, np.integer)) and not isinstance(axis, (int, np.integer)):
        values, axis = axis, values           # TF < 1.0 argument order
    ts = [_raw(v) for v in values]
    like = next((t for t in ts if t.is_floating_point()), ts[0])
    return T(torch.cat([t.to(like.dtype) if t.dtype != like.dtype and isinstance(v, (list, tuple, np.ndarray)) else t
                        for t, v in zip(ts, values)], dim=_int(axis)))


def stack(values, axis=0, name=None):
    ts = [_raw(v) for v in values]
    like = next((t for t in ts if t.is_floating_point()), ts[0])
    return T(torch.stack([t.to(like.dtype) if not isinstance(v, T) else t for t, v in zip(ts, values)], dim=_int(axis)))


def unstack(x, num=None, axis=0, name=None):
    return [T(t) for t in torch.unbind(_raw(x), dim=_int(axis))]


def transpose(x, perm=None, name=None):
    t = _raw(x)
    if perm is None:
        perm = list(reversed(range(t.dim())))
    return T(t.permute(*_ints(perm)))


def slice_(x, begin, size, name=None):
    t = _raw(x)
    idx = []
    for d, (b, s) in enumerate(zip(_ints(begin), _ints(size))):
        idx.append(slice(b, t.shape[d] if s == -1 else b + s))
    return T(t[tuple(idx)])


def gather(params, indices, validate_indices=None, name=None, axis=0):
    p = _raw(params)
    i = _raw(indices, dtype=torch.int64).long()
    ax = _int(axis)
    out = torch.index_select(p, ax, i.reshape(-1))
    return T(out.reshape(list(p.shape[:ax]) + list(i.shape) + list(p.shape[ax + 1:])))


def gather_nd(params, indices, name=None):
    p = _raw(params)
    i = _raw(indices, dtype=torch.int64).long()
    k = i.shape[-1]
    return T(p[tuple(i[..., j] for j in range(k))])


def one_hot(indices, depth, on_value=None, off_value=None, axis=None, dtype=None, name=None):
    i = _raw(indices, dtype=torch.int64).long()
    d = _dt(dtype) or (_raw(on_value).dtype if on_value is not None else _FLOAT[0])
    oh = torch.nn.functional.one_hot(i.clamp(0, _int(depth) - 1), _int(depth)).to(torch.bool)
    oh = oh & ((i >= 0) & (i < _int(depth))).unsqueeze(-1)      # out-of-range index -> all off (TF rule)
    on = torch.as_tensor(1 if on_value is None else float(_raw(on_value))).to(d)
    off = torch.as_tensor(0 if off_value is None else float(_raw(off_value))).to(d)
    out = torch.where(oh, on, off)
    if axis is not None and _int(axis) not in (-1, out.dim() - 1):
        out = out.movedim(-1, _int(axis))
    return T(out)


def where(condition, x=None, y=None, name=None):
    c = _raw(condition).bool()
    if x is None:
        return T(torch.nonzero(c).to(torch.int64))
    xt = _raw(x)
    yt = _raw(y, like=xt)
    if c.dim() == 1 and xt.dim() > 1:              # TF1: a vector condition selects whole rows
        c = c.reshape([-1] + [1] * (xt.dim() - 1))
    else:
        assert tuple(c.shape) == tuple(xt.shape), ('tf.where (TF1) does not broadcast', tuple(c.shape), tuple(xt.shape))
    return T(torch.where(c, xt, yt))


def dynamic_partition(data, partitions, num_partitions, name=None):
    d, p = _raw(data), _raw(partitions).long()
    flat_d = d.reshape([-1] + list(d.shape[p.dim():]))
    flat_p = p.reshape(-1)
    return [T(flat_d[flat_p == k]) for k in range(_int(num_partitions))]


def boolean_mask(x, mask, name=None):
    return T(_raw(x)[_raw(mask).bool()])


# ---- math -----------------------------------------------------------------------------------------------------------
def _red(fn):
    def f(x, axis=None, keep_dims=False, name=None, reduction_indices=None, keepdims=None):
        t = _raw(x)
        axis = axis if axis is not None else reduction_indices
        keep = bool(keep_dims if keepdims is None else keepdims)
        if axis is None:
            out = fn(t, tuple(range(t.dim())), False) if t.dim() else t
            return T(out.reshape([1] * t.dim()) if keep else out)
        return T(fn(t, tuple(a % t.dim() for a in _ints(axis)), keep))
    return f


def _mean(t, ax, keep):
    if not t.is_floating_point():          # integer mean truncates in TF
        return torch.div(t.sum(dim=ax, keepdim=keep), int(np.prod([t.shape[a] for a in ax])), rounding_mode='trunc')
    return t.mean(dim=ax, keepdim=keep)


reduce_sum = _red(lambda t, ax, k: t.sum(dim=ax, keepdim=k))
reduce_mean = _red(_mean)
reduce_max = _red(lambda t, ax, k: t.amax(dim=ax, keepdim=k))
reduce_min = _red(lambda t, ax, k: t.amin(dim=ax, keepdim=k))
reduce_all = _red(lambda t, ax, k: t.bool().all(dim=ax[0], keepdim=k) if len(ax) == 1 else t.bool().all())
reduce_any = _red(lambda t, ax, k: t.bool().any(dim=ax[0], keepdim=k) if len(ax) == 1 else t.bool().any())


def _prod(t, ax, keep):
    for a in sorted(ax, reverse=True):
        t = t.prod(dim=a, keepdim=keep)
    return t


reduce_prod = _red(_prod)


def norm(x, ord='euclidean', axis=None, keep_dims=False, name=None, keepdims=None):
    """tf.norm: sqrt(sum x^2) for ord 2 / 'euclidean' (vector norm over `axis`; no epsilon)."""
    assert ord in ('euclidean', 2, 2.0), ord
    t = _raw(x)
    keep = bool(keep_dims if keepdims is None else keepdims)
    if axis is None:
        return T(torch.sqrt((t * t).sum()))
    return T(torch.sqrt((t * t).sum(dim=tuple(a % t.dim() for a in _ints(axis)), keepdim=keep)))


def _un(fn):
    return lambda x, name=None: T(fn(_raw(x)))


abs_ = _un(torch.abs)
sin, cos, tan = _un(torch.sin), _un(torch.cos), _un(torch.tan)
exp, log, sqrt, square = _un(torch.exp), _un(torch.log), _un(torch.sqrt), _un(lambda t: t * t)
sigmoid, tanh, sign, floor, ceil, round_ = _un(torch.sigmoid), _un(torch.tanh), _un(torch.sign), _un(torch.floor), _un(torch.ceil), _un(torch.round)
negative = _un(torch.neg)
is_nan = _un(torch.isnan)
logical_not = _un(lambda t: ~t.bool())
atan = _un(torch.atan)


def _bin(fn):
    def f(x, y, name=None):
        a = _raw(x) if isinstance(x, T) or not isinstance(y, T) else None
        b = _raw(y, like=a) if a is not None else _raw(y)
        if a is None:
            a = _raw(x, like=b)
        return T(fn(a, b))
    return f


add, subtract, multiply = _bin(torch.add), _bin(torch.sub), _bin(torch.mul)
divide = div = truediv = _bin(lambda a, b: a / b if a.is_floating_point() else torch.div(a, b, rounding_mode='floor'))
maximum, minimum = _bin(torch.maximum), _bin(torch.minimum)
equal, not_equal = _bin(torch.eq), _bin(torch.ne)
greater, greater_equal, less, less_equal = _bin(torch.gt), _bin(torch.ge), _bin(torch.lt), _bin(torch.le)
logical_and, logical_or = _bin(lambda a, b: a.bool() & b.bool()), _bin(lambda a, b: a.bool() | b.bool())
pow_ = _bin(torch.pow)
atan2 = _bin(torch.atan2)
mod = floormod = _bin(torch.remainder)


def add_n(xs, name=None):
    out = _raw(xs[0])
    for v in xs[1:]:
        out = out + _raw(v, like=out)
    return T(out)


def clip_by_value(x, lo, hi, name=None):
    t = _raw(x)
    return T(torch.minimum(torch.maximum(t, _raw(lo, like=t)), _raw(hi, like=t)))


def matmul(a, b, transpose_a=False, transpose_b=False, name=None, **kw):
    x = _raw(a)
    y = _raw(b, like=x)
    if transpose_a:
        x = x.transpose(-1, -2)
    if transpose_b:
        y = y.transpose(-1, -2)
    return T(x @ y)


def argmax(x, axis=None, name=None, dimension=None, output_type=int64):
    ax = axis if axis is not None else dimension
    return T(torch.argmax(_raw(x), dim=_int(0 if ax is None else ax)).to(_dt(output_type)))


def argmin(x, axis=None, name=None, dimension=None, output_type=int64):
    ax = axis if axis is not None else dimension
    return T(torch.argmin(_raw(x), dim=_int(0 if ax is None else ax)).to(_dt(output_type)))


def cond(pred, true_fn=None, false_fn=None, name=None, fn1=None, fn2=None, strict=False):
    p = bool(_raw(pred)) if isinstance(pred, (T, torch.Tensor)) else bool(pred)
    return (true_fn or fn1)() if p else (false_fn or fn2)()


def _nest_map(fn, struct):
    if isinstance(struct, (list, tuple)):
        return type(struct)(_nest_map(fn, v) for v in struct)
    return fn(struct)


def _nest_flat(struct):
    if isinstance(struct, (list, tuple)):
        return [w for v in struct for w in _nest_flat(v)]
    return [struct]


def map_fn(fn, elems, dtype=None, parallel_iterations=None, back_prop=True, swap_memory=False, infer_shape=True, name=None):
    """tf.map_fn: fn over the leading axis of every tensor of the (possibly nested) structure `elems`; results stacked leaf by
    leaf in the structure fn returns (dtype only names that structure)."""
    elems = _nest_map(_w, elems)
    n = _nest_flat(elems)[0].t.shape[0]
    outs = [fn(_nest_map(lambda e: e[i], elems)) for i in range(n)]
    first = outs[0]
    if isinstance(first, (list, tuple)):
        flat = [_nest_flat(o) for o in outs]
        res = [T(torch.stack([_raw(f[k]) for f in flat])) for k in range(len(flat[0]))]
        it = iter(res)
        return _nest_map(lambda _: next(it), first)
    return T(torch.stack([_raw(o) for o in outs]))


def py_func(func, inp, Tout, stateful=True, name=None):
    args = [(_raw(v).detach().numpy()) for v in inp]
    out = func(*args)
    if isinstance(Tout, (list, tuple)):
        return [T(_raw(np.asarray(o), dtype=_dt(d))) for o, d in zip(out, Tout)]
    return T(_raw(np.asarray(out), dtype=_dt(Tout)))


def assert_greater(x, y, *a, **k):
    """Graph mode: the assert op is created but nothing depends on it, so a session never runs it (tf_util.py:504-505 would
    otherwise fail on every batch: height = top - bottom is negative in image coordinates)."""
    return None


def random_uniform(*a, **k):
    raise NotImplementedError('random ops are outside the pinned paths')


random_shuffle = random_normal = random_uniform


# ---- tf.nn ---------------------------------------------------------------------------------------------------------
def _softmax(logits, axis=-1, name=None, dim=None):
    return T(torch.softmax(_raw(logits), dim=_int(dim if dim is not None else axis)))


def _log_softmax(logits, axis=-1, name=None, dim=None):
    return T(torch.log_softmax(_raw(logits), dim=_int(dim if dim is not None else axis)))


def _sparse_xent(_sentinel=None, labels=None, logits=None, name=None):
    lg = _raw(logits)
    lb = _raw(labels).long()
    lsm = torch.log_softmax(lg, dim=-1)
    return T(-torch.gather(lsm, -1, lb.unsqueeze(-1)).squeeze(-1))


def _xent(_sentinel=None, labels=None, logits=None, dim=-1, name=None):
    lg = _raw(logits)
    return T(-(_raw(labels, like=lg) * torch.log_softmax(lg, dim=_int(dim))).sum(dim=_int(dim)))


def _sigmoid_xent(_sentinel=None, labels=None, logits=None, name=None):
    """max(x, 0) - x * z + log(1 + exp(-|x|)) (TF's documented stable form)."""
    x = _raw(logits)
    z = _raw(labels, like=x)
    return T(torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-x.abs())))


def _conv2d(input, filter, strides, padding, use_cudnn_on_gpu=True, data_format='NHWC', name=None, **kw):
    """NHWC, VALID (or SAME with a 1 x 1 kernel), as every conv2d call on the path is."""
    x, w = _raw(input), _raw(filter)
    assert data_format == 'NHWC'
    kh, kw_, cin, cout = w.shape
    assert padding == 'VALID' or (kh == 1 and kw_ == 1), padding
    out = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), stride=tuple(_ints(strides)[1:3]))
    return T(out.permute(0, 2, 3, 1))


def _bias_add(value, bias, data_format=None, name=None):
    assert data_format in (None, 'NHWC')
    return T(_raw(value) + _raw(bias))


def _max_pool(value, ksize, strides, padding, data_format='NHWC', name=None):
    x = _raw(value)
    k, s = _ints(ksize), _ints(strides)
    assert padding == 'VALID' and k[0] == 1 and k[3] == 1
    out = torch.nn.functional.max_pool2d(x.permute(0, 3, 1, 2), kernel_size=(k[1], k[2]), stride=(s[1], s[2]))
    return T(out.permute(0, 2, 3, 1))


def _dropout(x, keep_prob, noise_shape=None, seed=None, name=None):
    """x * keep_mask / keep_prob.  The keep mask is an explicit input, looked up by the enclosing scope name (the oracle's
    convention: VarStore.dropout_masks)."""
    t = _raw(x)
    full = '/'.join(p for p in STATE.scope if p)
    STATE.dropout_log.append(full)
    if full not in STATE.dropout_masks:
        raise KeyError('training-mode dropout needs an explicit keep mask for %s' % full)
    m = _raw(STATE.dropout_masks[full], like=t)
    assert m.numel() == t.numel(), (full, tuple(m.shape), tuple(t.shape))
    m = m.reshape(t.shape)          # masks are keyed (B, N, C); the graph's tensor is (B, N, 1, C)
    return T(t * m / float(_raw(keep_prob)))


def _moments(x, axes, shift=None, name=None, keep_dims=False):
    t = _raw(x)
    ax = tuple(_ints(axes))
    mean = t.mean(dim=ax, keepdim=True)
    var = ((t - mean) ** 2).mean(dim=ax, keepdim=True)
    if not keep_dims:
        mean, var = mean.reshape([s for i, s in enumerate(mean.shape) if i not in ax]), var.reshape([s for i, s in enumerate(var.shape) if i not in ax])
    return T(mean), T(var)


def _batch_normalization(x, mean, variance, offset, scale, variance_epsilon, name=None):
    t = _raw(x)
    inv = torch.rsqrt(_raw(variance) + variance_epsilon)
    if scale is not None:
        inv = inv * _raw(scale)
    out = (t - _raw(mean)) * inv
    return T(out + _raw(offset) if offset is not None else out)


def _l2_loss(t, name=None):
    x = _raw(t)
    return T((x * x).sum() / 2)


def _l2_normalize(x, axis=None, epsilon=1e-12, name=None, dim=None):
    t = _raw(x)
    ax = _ints(axis if axis is not None else dim)
    ss = (t * t).sum(dim=tuple(ax), keepdim=True)
    return T(t * torch.rsqrt(torch.clamp(ss, min=epsilon)))


def _relu(features, name=None):
    return T(torch.relu(_raw(features)))


def _tanh(x, name=None):
    return T(torch.tanh(_raw(x)))


def _leaky_relu(features, alpha=0.2, name=None):
    return T(torch.nn.functional.leaky_relu(_raw(features), alpha))


_relu.__name__, _tanh.__name__, _leaky_relu.__name__ = 'relu', 'tanh', 'leaky_relu'


nn = types.SimpleNamespace(
    relu=_relu, tanh=_tanh, sigmoid=_un(torch.sigmoid), leaky_relu=_leaky_relu,
    softmax=_softmax, log_softmax=_log_softmax,
    sparse_softmax_cross_entropy_with_logits=_sparse_xent, softmax_cross_entropy_with_logits=_xent,
    sigmoid_cross_entropy_with_logits=_sigmoid_xent,
    conv2d=_conv2d, bias_add=_bias_add, max_pool=_max_pool, dropout=_dropout, moments=_moments,
    batch_normalization=_batch_normalization, l2_loss=_l2_loss, l2_normalize=_l2_normalize)


# ---- tf.losses -----------------------------------------------------------------------------------------------------
class _Reduction(object):
    NONE = 'none'
    SUM = 'weighted_sum'
    MEAN = 'weighted_mean'
    SUM_BY_NONZERO_WEIGHTS = 'weighted_sum_by_nonzero_weights'
    SUM_OVER_BATCH_SIZE = 'weighted_sum_over_batch_size'


def _reduce_loss(l, weights, reduction):
    """tf.losses.compute_weighted_loss: elementwise * weights; default reduction SUM_BY_NONZERO_WEIGHTS = sum / number of
    elements with a non-zero (broadcast) weight."""
    w = torch.broadcast_to(_raw(weights, like=l), l.shape) if not isinstance(weights, (int, float)) else torch.full_like(l, float(weights))
    l = l * w
    if reduction == _Reduction.NONE:
        return T(l)
    if reduction == _Reduction.SUM:
        return T(l.sum())
    if reduction == _Reduction.MEAN:
        return T(l.sum() / w.sum())
    if reduction == _Reduction.SUM_OVER_BATCH_SIZE:
        return T(l.sum() / l.numel())
    nz = (w != 0).to(l.dtype).sum()
    return T(torch.where(nz > 0, l.sum() / torch.clamp(nz, min=1), torch.zeros_like(l.sum())))


def _huber_loss(labels, predictions, weights=1.0, delta=1.0, scope=None, loss_collection=None,
                reduction=_Reduction.SUM_BY_NONZERO_WEIGHTS):
    """0.5 e^2 for |e| <= delta, delta |e| - 0.5 delta^2 beyond (tf.losses.huber_loss: quadratic = min(|e|, delta),
    linear = |e| - quadratic, loss = 0.5 quadratic^2 + delta linear)."""
    p = _raw(predictions)
    y = _raw(labels, like=p)
    e = (p - y).abs()
    q = torch.clamp(e, max=delta)
    return _reduce_loss(0.5 * q * q + delta * (e - q), weights, reduction)


def _mse_loss(labels, predictions, weights=1.0, scope=None, loss_collection=None, reduction=_Reduction.SUM_BY_NONZERO_WEIGHTS):
    p = _raw(predictions)
    y = _raw(labels, like=p)
    return _reduce_loss((p - y) ** 2, weights, reduction)


losses = types.SimpleNamespace(Reduction=_Reduction, huber_loss=_huber_loss, mean_squared_error=_mse_loss)


# ---- tf.contrib.layers ---------------------------------------------------------------------------------------------
BN_EPS = 1e-3     # tf.contrib.layers.batch_norm default epsilon


def _contrib_batch_norm(inputs, decay=0.999, center=True, scale=False, epsilon=BN_EPS, activation_fn=None, param_initializers=None,
                        updates_collections='update_ops', is_training=True, reuse=None, variables_collections=None,
                        outputs_collections=None, trainable=True, data_format='NHWC', scope=None, **kw):
    """tf.contrib.layers.batch_norm with updates_collections=None (moving statistics updated in place by the forward pass).
    Variables scope/{beta,gamma,moving_mean,moving_variance}.  Training: batch mean and biased variance over all axes but
    the last normalise the batch; the moving variance receives the Bessel-corrected batch variance (what TF1's fused batch
    norm kernel returns for the update: variance * n / (n - 1)); moving <- decay * moving + (1 - decay) * batch."""
    assert data_format == 'NHWC' and center
    x = _raw(inputs)
    training = bool(_raw(is_training)) if isinstance(is_training, (T, torch.Tensor)) else bool(is_training)
    dec = float(_raw(decay))
    with variable_scope(scope or 'BatchNorm'):
        beta = get_variable('beta', [x.shape[-1]]).t
        gamma = get_variable('gamma', [x.shape[-1]]).t if scale else None
        mm = get_variable('moving_mean', [x.shape[-1]], trainable=False)
        mv = get_variable('moving_variance', [x.shape[-1]], trainable=False)
    if training:
        red = tuple(range(x.dim() - 1))
        n = int(np.prod([x.shape[d] for d in red]))
        mean = x.mean(dim=red)
        var = ((x - mean) ** 2).mean(dim=red)
        with torch.no_grad():
            mm.t.mul_(dec).add_((1 - dec) * mean.detach())
            mv.t.mul_(dec).add_((1 - dec) * var.detach() * (float(n) / max(n - 1, 1)))
    else:
        mean, var = mm.t, mv.t
    inv = torch.rsqrt(var + epsilon)
    if gamma is not None:
        inv = inv * gamma
    out = (x - mean) * inv + beta
    if activation_fn is not None:
        out = _raw(activation_fn(T(out)))
    return T(out)


contrib = types.SimpleNamespace(layers=types.SimpleNamespace(
    batch_norm=_contrib_batch_norm, xavier_initializer=lambda *a, **k: ('xavier',)))

summary = types.SimpleNamespace(scalar=lambda name, v, *a, **k: STATE.summaries.__setitem__(name, v), histogram=lambda *a, **k: None,
                                merge_all=lambda *a, **k: None, FileWriter=lambda *a, **k: None)


class _EMA(object):
    def __init__(self, decay, **kw):
        raise NotImplementedError('tf.train.ExponentialMovingAverage: only batch_norm_template_unused uses it (dead code)')


def _exp_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False, name=None):
    p = float(_raw(global_step)) / float(decay_steps)
    if staircase:
        p = np.floor(p)
    return T(torch.as_tensor(float(learning_rate) * float(decay_rate) ** p).to(_FLOAT[0]))


class _Saver(object):
    """tf.train.Saver: records the var_list it is built with ({checkpoint name: variable}); restore / save do nothing."""
    def __init__(self, var_list=None, **kw):
        self.var_list = var_list
        STATE.collections.setdefault('savers', []).append(self)

    def restore(self, sess, path):
        self.restored_from = path

    def save(self, *a, **k):
        return None


class _Optimizer(object):
    """tf.train.*Optimizer: minimize() only records what the step would differentiate (loss, var_list, global_step); the
    caller takes the gradients with torch.autograd."""
    def __init__(self, learning_rate=None, *a, **k):
        self.learning_rate = learning_rate

    def minimize(self, loss, global_step=None, var_list=None, **kw):
        op = types.SimpleNamespace(loss=loss, var_list=var_list, global_step=global_step, optimizer=self)
        STATE.collections.setdefault('train_ops', []).append(op)
        return op


train = types.SimpleNamespace(ExponentialMovingAverage=_EMA, exponential_decay=_exp_decay, Saver=_Saver,
                              AdamOptimizer=_Optimizer, MomentumOptimizer=_Optimizer, GradientDescentOptimizer=_Optimizer)


def install():
    """Builds and registers the `tensorflow` module object (a copy of this module's public names plus the names that would
    shadow Python builtins here: tf.range, tf.slice, tf.abs, tf.pow, tf.round, tf.bool)."""
    me = sys.modules[__name__]
    mod = types.ModuleType('tensorflow')
    for k, v in vars(me).items():
        if not k.startswith('__'):
            setattr(mod, k, v)
    for k, v in dict(range=range_, slice=slice_, abs=abs_, pow=pow_, round=round_, bool=bool_, Tensor=T).items():
        setattr(mod, k, v)
    mod.__version__ = '1.shim'
    sys.modules['tensorflow'] = mod
    return mod
