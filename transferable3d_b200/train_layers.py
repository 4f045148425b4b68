"""Building blocks of the training-mode graphs (train_boxpc.py, train_semisup_adv.py) on the B200:
flat parameter / gradient / Adam arenas keyed by TF variable names, one tf_util.conv2d / fully_connected
layer in training mode (x.W + b -> batch-statistics BN -> activation, tf_util.py:1258-1323, 1463-1499,
1645-1664) with its backward, and the eval-mode (frozen, BN folded) layer with dgrad only.
All arithmetic runs in libt3d_b200.so kernels (fp32 CUDA-core path).
"""
import numpy as np
import torch

from ._lib import ptr, stream, call
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


class TrainLayer(object):
    """One layer in training mode.  `params`: dict name -> fp32 device tensor for '<layer>/weights', '/biases',
    '/bn/gamma', '/bn/beta'; `moving`: dict for '/bn/moving_mean', '/bn/moving_variance' (updated in the forward
    pass, updates_collections=None); `grads`: dict of gradient views or None for a forward-only layer."""

    def __init__(self, name, kin, nout, bn, act, params, moving, grads=None):
        self.name, self.K, self.N, self.bn, self.act = name, kin, nout, bn, act
        self.params, self.moving, self.grads = params, moving, grads

    def p(self, suffix):
        return self.params[self.name + '/' + suffix]

    def W(self):
        return self.p('weights').view(self.K, self.N)

    def forward(self, x, bn_decay, y=None, keep=True):
        """x: (M,K) fp32.  `y`: pre-BN values computed by the caller (conv6 fold).  keep=False drops what only a
        backward pass would need."""
        M = x.shape[0] if x is not None else y.shape[0]
        dev = (x if x is not None else y).device
        if y is None:
            y = gemm(x, self.K, 1, self.W(), self.N, 1, M, self.N, self.K, bias=self.p('biases'))
        if keep:
            self.x, self.y = x, y
        if not self.bn:
            out = y
            if self.act != ACT_NONE:
                raise NotImplementedError('activation without batch norm is not on the hot path')
            self.out = out if keep else None
            return out
        s0 = torch.empty(self.N, device=dev)
        s1 = torch.empty(self.N, device=dev)
        # statistics of (y - row 0 of y): the per-column shift keeps the one-pass variance well conditioned
        call('t3d_colstats', ptr(y), None, ptr(y), None, None, ptr(s0), ptr(s1), M, self.N, 0, 0, stream())
        mean = torch.empty(self.N, device=dev)
        rstd = torch.empty(self.N, device=dev)
        call('t3d_bn_finalize', ptr(s0), ptr(s1), ptr(y), M, self.N, BN_EPS, float(bn_decay), ptr(mean), ptr(rstd),
             ptr(self.moving[self.name + '/bn/moving_mean']), ptr(self.moving[self.name + '/bn/moving_variance']), stream())
        out = torch.empty_like(y)
        call('t3d_bn_apply', ptr(y), ptr(mean), ptr(rstd), ptr(self.p('bn/gamma')), ptr(self.p('bn/beta')), ptr(out),
             M, self.N, self.act, stream())
        if keep:
            self.mean, self.rstd, self.out = mean, rstd, out
        return out

    def backward(self, dout, need_dx=True):
        """dout: gradient w.r.t. this layer's output (overwritten in place).  Returns dX or None.
        A layer whose variables are not in the gradient arena (frozen, but still in the training-mode graph: batch
        statistics) only propagates the input gradient -- TF back-propagates through frozen variables to their input."""
        M = self.y.shape[0]
        dev = dout.device
        g = self.grads
        trainable = g is not None and (self.name + '/weights') in g
        if not trainable and not need_dx:
            return None
        if self.bn:
            s1 = torch.empty(self.N, device=dev)
            s2 = torch.empty(self.N, device=dev)
            outp = ptr(self.out) if self.act != ACT_NONE else None
            call('t3d_colstats', ptr(dout), outp, ptr(self.y), ptr(self.mean), ptr(self.rstd), ptr(s1), ptr(s2), M, self.N, 1,
                 self.act, stream())
            if trainable:
                g[self.name + '/bn/beta'].copy_(s1)
                g[self.name + '/bn/gamma'].copy_(s2)
            call('t3d_bn_backward', ptr(dout), outp, ptr(self.y), ptr(self.mean), ptr(self.rstd), ptr(self.p('bn/gamma')),
                 ptr(s1), ptr(s2), M, self.N, self.act, stream())
        dy = dout
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
        # wgrad: dW[K,N] = X^T dY
        dW = g[self.name + '/weights'].view(self.K, self.N)
        call('t3d_gemm_f32', ptr(self.x), 1, self.K, ptr(dy), self.N, 1, ptr(dW), self.N, self.K, self.N, M,
             splitk_for(self.K, self.N, M), None, stream())
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

    def backward(self, dout, k_lo=0, k_hi=None):
        """dX[:, k_lo:k_hi] = (dout * act'(out)) . W[k_lo:k_hi, :]^T   (dout is overwritten)."""
        M = dout.shape[0]
        if self.act != ACT_NONE:
            call('t3d_act_bwd', ptr(dout), ptr(self.out), dout.numel(), self.act, stream())
        k_hi = self.K if k_hi is None else k_hi
        Wsub = self.Wf[k_lo:k_hi]                    # contiguous row block [k, N]
        return gemm(dout, self.N, 1, Wsub, 1, self.N, M, k_hi - k_lo, self.N)


def maxpool(x, B, N, C, rowmask=None):
    """max over the N rows of each group (+ arg-max); rowmask: pool x * rowmask[row] without materialising it."""
    pooled = torch.empty((B, C), dtype=torch.float32, device=x.device)
    arg = torch.empty((B, C), dtype=torch.int32, device=x.device)
    call('t3d_maxpool_masked_fwd', ptr(x), ptr(rowmask), B, N, C, ptr(pooled), ptr(arg), stream())
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
