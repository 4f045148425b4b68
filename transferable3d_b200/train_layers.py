"""Building blocks of the training-mode graphs (train_boxpc.py, train_semisup_adv.py) on the B200:
flat parameter / gradient / Adam arenas keyed by TF variable names, one tf_util.conv2d / fully_connected
layer in training mode (x.W + b -> batch-statistics BN -> activation, tf_util.py:1258-1323, 1463-1499,
1645-1664) with its backward, and the eval-mode (frozen, BN folded) layer with dgrad only.
All arithmetic runs in libt3d_b200.so kernels (fp32 CUDA-core path).
"""
import numpy as np
import torch

from ._lib import ptr, stream, call, gemm_workspace
from .constants import BN_EPS
from .weights import fold_bn

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_TANH = 0, 1, 2, 3


def gemm(A, sam, sak, Bm, sbk, sbn, M, N, K, bias=None, splitk=1, out=None):
    C = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=A.device)
    call('t3d_gemm_f32', ptr(A), sam, sak, ptr(Bm), sbk, sbn, ptr(C), N, M, N, K, splitk, ptr(bias), stream())
    return C


def splitk_for(M, N, K):
    tiles = ((M + 63) // 64) * ((N + 63) // 64)
    return int(max(1, min(K // 512, (1200 + tiles - 1) // tiles)))


class ParamArena(object):
    """Flat fp32 arenas (parameters, gradients, Adam m / v) with per-variable views; what the optimizer and the
    NCCL all-reduce see is one contiguous buffer (one all-reduce per step)."""

    def __init__(self, variables, names, device):
        sizes = [int(np.prod(np.asarray(variables[n]).shape)) for n in names]
        total = sum(sizes)
        self.names = list(names)
        self.flat_param = torch.empty(total, dtype=torch.float32, device=device)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=device)
        self.adam_m = torch.zeros(total, dtype=torch.float32, device=device)
        self.adam_v = torch.zeros(total, dtype=torch.float32, device=device)
        self.param, self.grad = {}, {}
        off = 0
        for n, sz in zip(names, sizes):
            self.param[n] = self.flat_param[off:off + sz]
            self.grad[n] = self.flat_grad[off:off + sz]
            self.param[n].copy_(torch.as_tensor(np.asarray(variables[n], dtype=np.float32).reshape(-1)))
            off += sz

    def adam_step(self, lr, t, world=1, pg=None, beta1=0.9, beta2=0.999, eps=1e-8):
        """tf.train.AdamOptimizer update after one all-reduce of the flat gradient arena."""
        from .dist_util import allreduce_flat
        world = allreduce_flat(self.flat_grad, pg) if world > 1 else 1
        lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
        call('t3d_adam', ptr(self.flat_param), ptr(self.flat_grad), ptr(self.adam_m), ptr(self.adam_v),
             self.flat_param.numel(), float(lr_t), beta1, beta2, eps, 1.0 / world, stream())


class Lazy(object):
    """Output of a lazy batch-norm layer: the pre-BN tensor y [M, C] plus the folded map of its batch norm; the value is
    relu(scale[c] * y + shift[c]).  Consumers that can apply the map in their loaders (TrainLayer.forward / wgrad on the
    tensor-core GEMM, maxpool) read y directly; anything else calls materialize() (one t3d_bn_apply pass, cached)."""

    def __init__(self, layer):
        self.layer, self.y, self.scale, self.shift = layer, layer.y, layer.a_scale, layer.a_shift
        self.shape, self.device = self.y.shape, self.y.device
        self._dense = None

    def materialize(self):
        if self._dense is None:
            l = self.layer
            out = torch.empty_like(self.y)
            call('t3d_bn_apply', ptr(self.y), ptr(l.mean), ptr(l.rstd), ptr(l.p('bn/gamma')), ptr(l.p('bn/beta')), ptr(out),
                 self.y.shape[0], self.y.shape[1], ACT_RELU, stream())
            self._dense = out
        return self._dense


def dense(x):
    return x.materialize() if isinstance(x, Lazy) else x


def bn_supported(M, N, K, kind):
    from ._lib import load
    return int(load().t3d_gemm_bn_supported(int(M), int(N), int(K), int(kind)))


class TrainLayer(object):
    """One layer in training mode.  `params`: dict name -> fp32 device tensor for '<layer>/weights', '/biases',
    '/bn/gamma', '/bn/beta'; `moving`: dict for '/bn/moving_mean', '/bn/moving_variance' (updated in the forward
    pass, updates_collections=None); `grads`: dict of gradient views or None for a forward-only layer.

    forward(..., lazy=True) (BN + ReLU layers) returns a Lazy instead of the activated tensor: the batch statistics come out
    of the GEMM epilogue (t3d_gemm_bn_f32), the BN map is applied by whoever consumes the output, and the backward pass
    recomputes the ReLU mask from y -- three of the five passes over each B*N x C activation of the forward pass and two of
    the nine of the backward pass disappear."""

    def __init__(self, name, kin, nout, bn, act, params, moving, grads=None):
        self.name, self.K, self.N, self.bn, self.act = name, kin, nout, bn, act
        self.params, self.moving, self.grads = params, moving, grads
        self.lazy_out = False

    def p(self, suffix):
        return self.params[self.name + '/' + suffix]

    def W(self):
        return self.p('weights').view(self.K, self.N)

    def forward(self, x, bn_decay, y=None, keep=True, lazy=False):
        """x: (M,K) fp32 tensor or Lazy.  `y`: pre-BN values computed by the caller (conv6 fold).  keep=False drops what only
        a backward pass would need."""
        M = x.shape[0] if x is not None else y.shape[0]
        dev = (x if x is not None else y).device
        lazy = bool(lazy and self.bn and self.act == ACT_RELU)
        E = lambda: torch.empty(self.N, device=dev)
        fused = False
        if y is None:
            kind = bn_supported(M, self.N, self.K, 0) if self.bn else 0

            if isinstance(x, Lazy) and kind != 1:
                x = x.materialize()
            if kind:
                # statistics in the GEMM epilogue, shifted by row 0 of the output (t3d_row0) as t3d_colstats does
                xl = x if isinstance(x, Lazy) else None
                xa = xl.y if xl is not None else x
                sc, sh = (ptr(xl.scale), ptr(xl.shift)) if xl is not None else (None, None)
                y0, s0, s1 = E(), E(), E()
                call('t3d_row0', ptr(xa), sc, sh, ptr(self.W()), self.N, ptr(self.p('biases')), self.K, self.N, ptr(y0), stream())
                y = torch.empty((M, self.N), dtype=torch.float32, device=dev)
                ws = gemm_workspace()
                call('t3d_gemm_bn_f32', ptr(xa), self.K, 1, sc, sh, ptr(self.W()), self.N, 1, ptr(y), self.N, M, self.N, self.K, 1,
                     ptr(self.p('biases')), ptr(s0), ptr(s1), ptr(y0), ptr(ws), ws.numel(), stream())
                fused = True
            else:
                y = gemm(x, self.K, 1, self.W(), self.N, 1, M, self.N, self.K, bias=self.p('biases'))
        if keep:
            self.x, self.y = x, y
        if not self.bn:
            out = y
            if self.act != ACT_NONE:
                raise NotImplementedError('activation without batch norm is not on the hot path')
            self.out = out if keep else None
            return out
        if not fused:
            s0, s1, y0 = E(), E(), y
            # statistics of (y - row 0 of y): the per-column shift keeps the one-pass variance well conditioned
            call('t3d_colstats', ptr(y), None, ptr(y), None, None, ptr(s0), ptr(s1), M, self.N, 0, 0, stream())
        mean, rstd = E(), E()
        mm, mv = self.moving[self.name + '/bn/moving_mean'], self.moving[self.name + '/bn/moving_variance']
        if lazy:
            a_scale, a_shift = E(), E()
            call('t3d_bn_finalize_affine', ptr(s0), ptr(s1), ptr(y0), M, self.N, BN_EPS, float(bn_decay), ptr(self.p('bn/gamma')),
                 ptr(self.p('bn/beta')), ptr(mean), ptr(rstd), ptr(a_scale), ptr(a_shift), ptr(mm), ptr(mv), stream())
            self.y, self.mean, self.rstd, self.a_scale, self.a_shift = y, mean, rstd, a_scale, a_shift
            res = Lazy(self)
            self.lazy_out = True
            self.out = None
            if not keep:
                self.y = None
            return res
        call('t3d_bn_finalize', ptr(s0), ptr(s1), ptr(y0), M, self.N, BN_EPS, float(bn_decay), ptr(mean), ptr(rstd), ptr(mm), ptr(mv),
             stream())
        out = torch.empty_like(y)
        call('t3d_bn_apply', ptr(y), ptr(mean), ptr(rstd), ptr(self.p('bn/gamma')), ptr(self.p('bn/beta')), ptr(out),
             M, self.N, self.act, stream())
        self.lazy_out = False
        if keep:
            self.mean, self.rstd, self.out = mean, rstd, out
        return out

    def forward_pooled(self, x, bn_decay, B, N):
        """Forward-only BN + ReLU layer whose output only feeds the max-pool over the N rows of each of the B groups: returns the
        pooled activations [B, C].  On the tensor-core path the statistics and the per-group max / min of the pre-BN output
        come out of the GEMM epilogue and the B*N x C output is never written (t3d_gemm_bn_pool_f32); otherwise forward +
        maxpool."""
        M = B * N
        ok = self.bn and self.act == ACT_RELU and N % 128 == 0 and bn_supported(M, self.N, self.K, 0) == 1
        if not ok:
            pooled, _ = maxpool(self.forward(x, bn_decay, keep=False, lazy=True), B, N, self.N)
            return pooled
        dev = x.device
        E = lambda: torch.empty(self.N, device=dev)
        xl = x if isinstance(x, Lazy) else None
        xa = xl.y if xl is not None else x
        sc, sh = (ptr(xl.scale), ptr(xl.shift)) if xl is not None else (None, None)
        y0, s0, s1, mean, rstd, a_scale, a_shift = E(), E(), E(), E(), E(), E(), E()
        call('t3d_row0', ptr(xa), sc, sh, ptr(self.W()), self.N, ptr(self.p('biases')), self.K, self.N, ptr(y0), stream())
        kmax = torch.empty((B, self.N), dtype=torch.int32, device=dev)
        kmin = torch.empty((B, self.N), dtype=torch.int32, device=dev)
        ws = gemm_workspace()
        call('t3d_gemm_bn_pool_f32', ptr(xa), self.K, sc, sh, ptr(self.W()), self.N, M, self.N, self.K, ptr(self.p('biases')), ptr(s0),
             ptr(s1), ptr(y0), N, ptr(kmax), ptr(kmin), ptr(ws), ws.numel(), stream())
        mm, mv = self.moving[self.name + '/bn/moving_mean'], self.moving[self.name + '/bn/moving_variance']
        call('t3d_bn_finalize_affine', ptr(s0), ptr(s1), ptr(y0), M, self.N, BN_EPS, float(bn_decay), ptr(self.p('bn/gamma')),
             ptr(self.p('bn/beta')), ptr(mean), ptr(rstd), ptr(a_scale), ptr(a_shift), ptr(mm), ptr(mv), stream())
        pooled = torch.empty((B, self.N), dtype=torch.float32, device=dev)
        call('t3d_pool_bn_finish', ptr(kmax), ptr(kmin), ptr(a_scale), ptr(a_shift), B, self.N, ptr(pooled), stream())
        return pooled

    def backward(self, dout, need_dx=True):
        """dout: gradient w.r.t. this layer's output (overwritten in place).  Returns dX or None.
        A layer whose variables are not in the gradient arena (frozen, but still in the training-mode graph: batch
        statistics) only propagates the input gradient -- TF back-propagates through frozen variables to their input."""
        g = self.grads
        trainable = g is not None and (self.name + '/weights') in g
        if not trainable and not need_dx:
            return None
        return self._backward_gemms(self._backward_bn(dout), trainable, need_dx)

    def _backward_bn(self, dout):
        """dout (gradient w.r.t. the layer output) -> gradient w.r.t. the pre-BN values, in place; d beta / d gamma."""
        if not self.bn:
            return dout
        M = self.y.shape[0]
        dev = dout.device
        g = self.grads
        trainable = g is not None and (self.name + '/weights') in g
        s1 = torch.empty(self.N, device=dev)
        s2 = torch.empty(self.N, device=dev)
        if self.lazy_out:
            call('t3d_colstats_lazy', ptr(dout), ptr(self.y), ptr(self.mean), ptr(self.rstd), ptr(self.a_scale), ptr(self.a_shift),
                 ptr(s1), ptr(s2), M, self.N, stream())
            call('t3d_bn_backward_lazy', ptr(dout), ptr(self.y), ptr(self.mean), ptr(self.rstd), ptr(self.p('bn/gamma')),
                 ptr(self.a_scale), ptr(self.a_shift), ptr(s1), ptr(s2), M, self.N, stream())
        else:
            outp = ptr(self.out) if self.act != ACT_NONE else None
            call('t3d_colstats', ptr(dout), outp, ptr(self.y), ptr(self.mean), ptr(self.rstd), ptr(s1), ptr(s2), M, self.N, 1,
                 self.act, stream())
            call('t3d_bn_backward', ptr(dout), outp, ptr(self.y), ptr(self.mean), ptr(self.rstd), ptr(self.p('bn/gamma')),
                 ptr(s1), ptr(s2), M, self.N, self.act, stream())
        if trainable:
            g[self.name + '/bn/beta'].copy_(s1)
            g[self.name + '/bn/gamma'].copy_(s2)
        return dout

    def backward_bn_only(self, dout):
        """BN (+ activation) part of backward() for a layer whose pre-BN values were computed by the caller (forward(None, y=...)):
        returns dY (dout, overwritten) and leaves d gamma / d beta in the gradient arena; the caller owns the GEMMs."""
        return self._backward_bn(dout)

    def backward_pooled(self, gpool, arg, B, N, rowmask=None, need_dx=True):
        """Backward of this (lazy BN + ReLU) layer when its output only feeds a max-pool over the N rows of each group
        (optionally through `* rowmask`): gpool [B, C] is the gradient w.r.t. the pooled features, `arg` the arg-max rows.
        The pooled gradient is never scattered into a dense B*N x C tensor (t3d_pool_bn_backward)."""
        assert self.bn and self.lazy_out
        dev = gpool.device

        g = self.grads
        trainable = g is not None and (self.name + '/weights') in g
        if not trainable and not need_dx:
            return None
        s1 = torch.empty(self.N, device=dev)
        s2 = torch.empty(self.N, device=dev)
        dY = torch.empty((B * N, self.N), dtype=torch.float32, device=dev)
        call('t3d_pool_bn_backward', ptr(gpool), ptr(arg), ptr(rowmask), ptr(self.y), ptr(self.mean), ptr(self.rstd),
             ptr(self.p('bn/gamma')), ptr(self.a_scale), ptr(self.a_shift), B, N, self.N, ptr(s1), ptr(s2), ptr(dY), stream())
        if trainable:
            g[self.name + '/bn/beta'].copy_(s1)
            g[self.name + '/bn/gamma'].copy_(s2)
        return self._backward_gemms(dY, trainable, need_dx)

    def _backward_gemms(self, dy, trainable, need_dx):
        M = dy.shape[0]
        dev = dy.device
        g = self.grads
        if not trainable:
            return gemm(dy, self.N, 1, self.W(), 1, self.N, M, self.K, self.N)
        if self.bn:
            # a bias in front of a batch norm cancels in (y - mean): its gradient, the column sum of the BN input gradient,
            # is analytically zero (what TF accumulates there is rounding noise) -- no pass over dY for it
            g[self.name + '/biases'].zero_()
        else:
            bs = torch.empty(self.N, device=dev)
            junk = torch.empty(self.N, device=dev)
            call('t3d_colstats', ptr(dy), None, None, None, None, ptr(bs), ptr(junk), M, self.N, 0, 0, stream())
            g[self.name + '/biases'].copy_(bs)
        # wgrad: dW[K,N] = X^T dY   (X lazy: the BN map of the previous layer is applied by the loader, per row of X^T)
        dW = g[self.name + '/weights'].view(self.K, self.N)
        x = self.x
        sk = splitk_for(self.K, self.N, M)
        if isinstance(x, Lazy) and bn_supported(self.K, self.N, M, 1):
            ws = gemm_workspace()
            call('t3d_gemm_bn_f32', ptr(x.y), 1, self.K, ptr(x.scale), ptr(x.shift), ptr(dy), self.N, 1, ptr(dW), self.N, self.K, self.N,
                 M, sk, None, None, None, None, ptr(ws), ws.numel(), stream())
        else:
            x = dense(x)
            call('t3d_gemm_f32', ptr(x), 1, self.K, ptr(dy), self.N, 1, ptr(dW), self.N, self.K, self.N, M, sk, None, stream())
        if not need_dx:
            return None
        # dgrad: dX[M,K] = dY W^T
        return gemm(dy, self.N, 1, self.W(), 1, self.N, M, self.K, self.N)


class EvalLayer(object):
    """One frozen layer in eval mode (moving-statistics BN folded into W, b): forward keeps the output so that the
    input gradient can be propagated (dgrad only; the frozen BoxPC branch of train_semisup_adv.py:337-345,364-388)."""

    def __init__(self, variables, layer, act, device):
        w, b = fold_bn(variables, layer)
        self.Wf = torch.from_numpy(w).to(device).contiguous()
        self.bf = torch.from_numpy(b).to(device).contiguous()
        self.K, self.N = self.Wf.shape
        self.act = act

    def forward(self, x):
        M = x.shape[0]
        y = torch.empty((M, self.N), dtype=torch.float32, device=x.device)
        call('t3d_linear_f32', ptr(x), self.K, ptr(self.Wf), self.N, ptr(self.bf), None, 0, ptr(y), self.N, M, self.K, self.N,
             self.act, None, None, stream())
        self.out = y
        return y

    def backward(self, dout, k_lo=0, k_hi=None, out=None):
        """dX[:, k_lo:k_hi] = (dout * act'(out)) . W[k_lo:k_hi, :]^T   (dout is overwritten).  `out`: this layer's forward
        output at the rows of dout when the backward pass runs on a row subset (pool_rows / gather_rows)."""
        M = dout.shape[0]
        out = self.out if out is None else out
        assert out.shape == dout.shape
        if self.act != ACT_NONE:
            call('t3d_act_bwd', ptr(dout), ptr(out), dout.numel(), self.act, stream())
        k_hi = self.K if k_hi is None else k_hi
        Wsub = self.Wf[k_lo:k_hi]                    # contiguous row block [k, N]
        return gemm(dout, self.N, 1, Wsub, 1, self.N, M, k_hi - k_lo, self.N)


def maxpool(x, B, N, C, rowmask=None):
    """max over the N rows of each group (+ arg-max); rowmask: pool x * rowmask[row] without materialising it.
    x may be a Lazy: the BN map + ReLU is applied while pooling."""
    pooled = torch.empty((B, C), dtype=torch.float32, device=x.device)
    arg = torch.empty((B, C), dtype=torch.int32, device=x.device)
    keys = torch.empty((B, C), dtype=torch.int64, device=x.device)          # scratch of the row-split kernel

    if isinstance(x, Lazy):
        call('t3d_maxpool_fwd_ws', ptr(x.y), ptr(x.scale), ptr(x.shift), ptr(rowmask), B, N, C, ptr(pooled), ptr(arg), ptr(keys), stream())
    else:
        call('t3d_maxpool_fwd_ws', ptr(x), None, None, ptr(rowmask), B, N, C, ptr(pooled), ptr(arg), ptr(keys), stream())
    return pooled, arg


def maxpool_bwd(g, arg, B, N, C, rowmask=None):
    dx = torch.empty((B * N, C), dtype=torch.float32, device=g.device)
    call('t3d_maxpool_masked_bwd', ptr(g), ptr(arg), ptr(rowmask), B, N, C, ptr(dx), stream())
    return dx


def rowmask_mul(x, rowmask, inplace=False):
    out = x if inplace else torch.empty_like(x)
    call('t3d_rowmask_mul', ptr(x), ptr(rowmask), ptr(out), x.shape[0], x.shape[1], stream())
    return out


def dropout(x, keep_mask, keep_prob):
    """tf.nn.dropout with an explicit keep mask (forward and backward are the same map)."""
    out = torch.empty_like(x)
    call('t3d_scale_mask', ptr(x), ptr(keep_mask), 1.0 / keep_prob, ptr(out), x.numel(), stream())
    return out


def pool_rows(arg, B, N, C):
    """Distinct arg-max rows of every frustum: rows [B,S] (ascending, -1 padded), slot [B,C], count [B]; S = min(C, N)."""
    S = min(C, N)
    rows = torch.empty((B, S), dtype=torch.int32, device=arg.device)
    slot = torch.empty((B, C), dtype=torch.int32, device=arg.device)
    count = torch.empty((B,), dtype=torch.int32, device=arg.device)
    call('t3d_pool_rows', ptr(arg), B, N, C, S, ptr(rows), ptr(slot), ptr(count), stream())
    return rows, slot, count, S


def gather_rows(src, rows, B, N, S):
    """src [B*N, C] -> [B*S, C]: the rows listed per frustum (zeros where -1)."""
    C = src.shape[-1]
    dst = torch.empty((B * S, C), dtype=torch.float32, device=src.device)
    call('t3d_gather_rows', ptr(src), ptr(rows), B, N, S, C, ptr(dst), stream())
    return dst


def scatter_pool_grad(g, slot, B, C, S):
    dst = torch.empty((B * S, C), dtype=torch.float32, device=g.device)
    call('t3d_scatter_pool_grad', ptr(g), ptr(slot), B, C, S, ptr(dst), stream())
    return dst
