"""-m gpu: the tcgen05 "bf16 x 3" fp32 GEMM (csrc/xgemm.cuh) behind t3d_gemm_f32 / t3d_linear_f32, through the C ABI.

The checker is a float64 matmul of the same fp32 inputs (test infrastructure only).  Bars:
  * error of the tensor-core engine <= a few fp32 ulps of the row's |a|.|b| scale -- the same bar the CUDA-core SGEMM
    is held to, and both engines are measured side by side;
  * every operand layout of the training step (forward: A unit-k / B unit-n, dgrad: both unit-k, wgrad: both unit-row
    with split-K), ragged M / N / K, the K tail, bias, split-K reductions;
  * t3d_linear_f32 epilogues: bias, per-group bias, activations, row mask, group max with and without Y.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rt():
    from transferable3d_b200 import runtime as rt
    return rt


def _gemm(A, sam, sak, Bm, sbk, sbn, M, N, K, bias=None, splitk=1):
    from transferable3d_b200.train_layers import gemm
    return gemm(A, sam, sak, Bm, sbk, sbn, M, N, K, bias=bias, splitk=splitk)


def _operands(M, N, K, a_unit_k, b_unit_k, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    # wide dynamic range inside every dot product: exercises all three bf16 pieces
    a = torch.randn(M, K, generator=g, device='cuda') * torch.exp(2.0 * torch.randn(M, K, generator=g, device='cuda'))
    b = torch.randn(K, N, generator=g, device='cuda') * torch.exp(2.0 * torch.randn(K, N, generator=g, device='cuda'))
    A = a.contiguous() if a_unit_k else a.t().contiguous()            # [M,K] or [K,M]
    Bm = b.t().contiguous() if b_unit_k else b.contiguous()           # [N,K] or [K,N]
    sam, sak = (K, 1) if a_unit_k else (1, M)
    sbk, sbn = (1, K) if b_unit_k else (N, 1)
    return a, b, A, sam, sak, Bm, sbk, sbn


def _ulp_err(C, a, b):
    ref = a.double() @ b.double()
    scale = a.double().abs() @ b.double().abs()          # sum_k |a||b|: the natural error scale of a dot product
    return float(((C.double() - ref).abs() / scale).max()) / 2.0 ** -24


SHAPES = [(256, 128, 128), (384, 256, 96), (1000, 200, 333), (128, 64, 32), (130, 515, 70), (4096, 512, 256), (4500, 200, 333),
          (64, 512, 8192), (100, 130, 5000)]      # M < 128 with a long reduction: the weight gradient of a 64-channel layer


@pytest.mark.parametrize('a_unit_k,b_unit_k', [(True, False), (True, True), (False, False), (False, True)])
@pytest.mark.parametrize('M,N,K', SHAPES)
def test_gemm_layouts_fp32_accurate(M, N, K, a_unit_k, b_unit_k):
    rt = _rt()
    a, b, A, sam, sak, Bm, sbk, sbn = _operands(M, N, K, a_unit_k, b_unit_k, seed=M + N + K)
    bias = torch.randn(N, device='cuda')
    with rt.f32_engine('tc'):
        C = _gemm(A, sam, sak, Bm, sbk, sbn, M, N, K)
        Cb = _gemm(A, sam, sak, Bm, sbk, sbn, M, N, K, bias=bias)
    with rt.f32_engine('simt'):
        C0 = _gemm(A, sam, sak, Bm, sbk, sbn, M, N, K)
    torch.cuda.synchronize()
    e_tc, e_simt = _ulp_err(C, a, b), _ulp_err(C0, a, b)
    # the bar is the CUDA-core fp32 SGEMM itself (K sequential roundings): measured 12.8 ulp of sum|a||b| at K = 128 on the
    # worst of 32768 outputs, against 13.7 for the split
    # (reductions longer than one 2048-deep chunk add the tensor core's truncating accumulation, ~0.5 ulp per K = 16 step inside
    # a chunk: 68 ulp against 45 for the SGEMM at K = 5000, held to 2 x there)
    assert e_tc < max(4.0, (1.5 if K <= 2048 else 2.0) * e_simt), 'tcgen05 bf16x3: %.2f ulp of sum|a||b| (CUDA-core SGEMM: %.2f)' % (e_tc, e_simt)
    if K <= 2048:
        assert torch.equal(Cb, C + bias)
    else:      # K ranges beyond 2048 are chunked and summed with atomics: the order of the partial sums varies from call to call
        assert float(((Cb - bias - C).abs() / (a.double().abs() @ b.double().abs()).float()).max()) < 1e-6


@pytest.mark.parametrize('splitk', [2, 7, 33])
def test_wgrad_splitk(splitk):
    rt = _rt()
    M, N, K = 128, 256, 20000         # dW[Cin,Cout] = X^T dY over 20000 rows
    a, b, A, sam, sak, Bm, sbk, sbn = _operands(M, N, K, False, False, seed=splitk)
    with rt.f32_engine('tc'):
        C = _gemm(A, sam, sak, Bm, sbk, sbn, M, N, K, splitk=splitk)
        C1 = _gemm(A, sam, sak, Bm, sbk, sbn, M, N, K, splitk=1)
    torch.cuda.synchronize()
    with rt.f32_engine('simt'):
        C0 = _gemm(A, sam, sak, Bm, sbk, sbn, M, N, K, splitk=splitk)
    e, e1, e0 = _ulp_err(C, a, b), _ulp_err(C1, a, b), _ulp_err(C0, a, b)
    # K chunks are capped at 2048 inside the call (the accumulator truncates: xgemm.cuh), so splitk = 1 is a 10-way split
    assert e < 128.0 and e1 < 128.0, 'ulp of sum|a||b|: tc %.1f, tc splitk=1 %.1f, CUDA cores %.1f' % (e, e1, e0)


def test_exact_on_bf16_representable_inputs():
    """Inputs with <= 8 significant bits and small integer values: every partial product and sum is exact."""
    rt = _rt()
    g = torch.Generator(device='cuda').manual_seed(3)
    a = torch.randint(-8, 9, (512, 192), generator=g, device='cuda').float()
    b = torch.randint(-8, 9, (192, 128), generator=g, device='cuda').float()
    with rt.f32_engine('tc'):
        C = _gemm(a.contiguous(), 192, 1, b.contiguous(), 128, 1, 512, 128, 192)
    assert torch.equal(C, a @ b)


def test_split_is_exact_for_full_mantissas():
    """24-bit mantissas times a power of two: the product needs all three pieces of a."""
    rt = _rt()
    g = torch.Generator(device='cuda').manual_seed(4)
    a = (torch.randint(2 ** 23, 2 ** 24, (256, 64), generator=g, device='cuda').float()) * 2.0 ** -20
    b = torch.zeros(64, 128, device='cuda')
    b[torch.arange(64), torch.arange(64)] = 4.0          # C[:, j] = 4 a[:, j] for j < 64
    with rt.f32_engine('tc'):
        C = _gemm(a.contiguous(), 64, 1, b.contiguous(), 128, 1, 256, 128, 64)
    assert torch.equal(C[:, :64], 4.0 * a) and float(C[:, 64:].abs().max()) == 0.0


@pytest.mark.parametrize('G', [6, 20])                 # M = 1536: B split per tile; M = 5120: pre-split B (workspace path)
@pytest.mark.parametrize('act', [None, 'relu', 'leaky_relu', 'tanh'])
def test_linear_epilogues(act, G):
    rt = _rt()
    R, K, N = 256, 128, 200                           # G groups of 256 rows
    M = G * R
    g = torch.Generator(device='cuda').manual_seed(11)
    x = torch.randn(M, K, generator=g, device='cuda')
    w = torch.randn(K, N, generator=g, device='cuda') * 0.1
    b = torch.randn(N, generator=g, device='cuda')
    gb = torch.randn(G, N, generator=g, device='cuda')
    rm = (torch.rand(M, generator=g, device='cuda') < 0.5).float()
    pre = (x.double() @ w.double() + b.double()).view(G, R, N) + gb.double()[:, None, :]
    f = {None: lambda v: v, 'relu': torch.relu, 'leaky_relu': lambda v: torch.where(v > 0, v, 0.2 * v), 'tanh': torch.tanh}[act]
    ref = (f(pre) * rm.double().view(G, R, 1)).view(M, N)
    with rt.f32_engine('tc'):
        y, _ = rt.linear(x, w, b, act, gbias=gb, rows_per_group=R, rowmask=rm)
    assert float((y.double() - ref).abs().max()) < 2e-5
    if act == 'relu':
        with rt.f32_engine('tc'):
            y2, gmax = rt.linear(x, w, b, act, gbias=gb, rows_per_group=R, rowmask=rm, gmax_groups=G)
            _, gmax_only = rt.linear(x, w, b, act, gbias=gb, rows_per_group=R, rowmask=rm, gmax_groups=G, want_y=False)
            # groups that do not align with the 128-row tiles take the per-element path
            _, gmax_ragged = rt.linear(x[:5 * 300], w, b, act, rows_per_group=300, gmax_groups=5, want_y=False)
        assert torch.equal(y2, y)
        assert torch.equal(gmax, y.view(G, R, N).max(dim=1).values.clamp_min(0.0))
        assert torch.equal(gmax_only, gmax)
        ref_r = torch.relu(x[:1500].double() @ w.double() + b.double()).view(5, 300, N).max(dim=1).values
        assert float((gmax_ragged.double() - ref_r).abs().max()) < 2e-5


def test_engines_agree_on_a_training_layer():
    """forward / dgrad / wgrad of one 128 -> 256 layer over 32768 rows: tensor-core engine vs CUDA-core engine."""
    rt = _rt()
    M, K, N = 32768, 128, 256
    g = torch.Generator(device='cuda').manual_seed(5)
    x = torch.randn(M, K, generator=g, device='cuda')
    w = torch.randn(K, N, generator=g, device='cuda') * 0.1
    dy = torch.randn(M, N, generator=g, device='cuda')
    out = {}
    for eng in ('tc', 'simt'):
        with rt.f32_engine(eng):
            y = _gemm(x, K, 1, w, N, 1, M, N, K)
            dx = _gemm(dy, N, 1, w, 1, N, M, K, N)
            dw = _gemm(x, 1, K, dy, N, 1, K, N, M, splitk=37)
        out[eng] = (y, dx, dw)
    ref = (x.double() @ w.double(), dy.double() @ w.double().t(), x.double().t() @ dy.double())
    for i, name in enumerate(('forward', 'dgrad', 'wgrad')):
        scale = float(ref[i].abs().mean())
        e_tc = float((out['tc'][i].double() - ref[i]).abs().max()) / scale
        e_simt = float((out['simt'][i].double() - ref[i]).abs().max()) / scale
        assert e_tc < 1e-5, '%s: tcgen05 engine %.3g of scale (CUDA cores %.3g)' % (name, e_tc, e_simt)
        assert e_tc < 4.0 * e_simt + 1e-6, '%s: tcgen05 engine %.3g vs CUDA cores %.3g' % (name, e_tc, e_simt)


@pytest.mark.parametrize('M', [1024, 8192])            # without / with the pre-split workspace path
def test_bf16_engine_is_a_bf16_gemm(M):
    """Engine 'bf16': operands rounded to nearest bf16, fp32 accumulation -- equals a float64 matmul of the rounded inputs
    up to accumulation order."""
    rt = _rt()
    K, N = 256, 384
    g = torch.Generator(device='cuda').manual_seed(9)
    x = torch.randn(M, K, generator=g, device='cuda')
    w = torch.randn(K, N, generator=g, device='cuda') * 0.1
    with rt.f32_engine('bf16'):
        y = _gemm(x, K, 1, w, N, 1, M, N, K)
        y2, _ = rt.linear(x, w, None, 'relu')
    ref = x.bfloat16().double() @ w.bfloat16().double()
    assert float((y.double() - ref).abs().max()) < 2e-5
    assert float((y2.double() - torch.relu(ref)).abs().max()) < 2e-5
    full = x.double() @ w.double()
    assert float((y.double() - full).abs().max()) / float(full.abs().mean()) < 5e-2       # bf16-class accuracy
    assert rt.get_f32_engine() == 'tc'


@pytest.mark.parametrize('Kin', [3, 12])
def test_skinny_first_layer_shapes(Kin):
    """skinny_gemm.cuh: forward with K = 3 / 12, input gradient with N = 3 / 12, weight gradient with M = 3 / 12 over
    B*N rows -- the HBM-bound first-layer shapes of the point networks -- against a float64 matmul."""
    rt = _rt()
    rows, N = 70000, 128
    g = torch.Generator(device='cuda').manual_seed(Kin)
    x = torch.randn(rows, Kin, generator=g, device='cuda')
    w = torch.randn(Kin, N, generator=g, device='cuda') * 0.3
    b = torch.randn(N, generator=g, device='cuda')
    dy = torch.randn(rows, N, generator=g, device='cuda')
    y = _gemm(x, Kin, 1, w, N, 1, rows, N, Kin, bias=b)                       # forward
    dx = _gemm(dy, N, 1, w, 1, N, rows, Kin, N)                               # dgrad: dY W^T
    dw = _gemm(x, 1, Kin, dy, N, 1, Kin, N, rows, splitk=64)                  # wgrad: X^T dY
    dw1 = _gemm(x, 1, Kin, dy, N, 1, Kin, N, rows, splitk=1)
    torch.cuda.synchronize()
    assert float((y.double() - (x.double() @ w.double() + b.double())).abs().max()) < 1e-5
    assert float((dx.double() - dy.double() @ w.double().t()).abs().max()) < 2e-5
    ref = x.double().t() @ dy.double()
    assert float((dw.double() - ref).abs().max()) < 2e-4 * float(ref.abs().max())
    assert float((dw1.double() - ref).abs().max()) < 2e-4 * float(ref.abs().max())


def test_skinny_linear_first_and_last_layers():
    """t3d_linear_f32 on the HBM-bound layers of the fp32-mode pipeline: conv1 (K = 6 / 3 inputs, bias + ReLU) and conv10
    (N = 2 outputs, bias only), M >= 4096 rows."""
    rt = _rt()
    g = torch.Generator(device='cuda').manual_seed(21)
    M = 9000
    for K, N, act in ((6, 64, 'relu'), (3, 128, 'relu'), (12, 128, None), (128, 2, None), (256, 12, None)):
        x = torch.randn(M, K, generator=g, device='cuda')
        w = torch.randn(K, N, generator=g, device='cuda') * 0.2
        b = torch.randn(N, generator=g, device='cuda')
        y, _ = rt.linear(x, w, b, act)
        ref = x.double() @ w.double() + b.double()
        if act == 'relu':
            ref = torch.relu(ref)
        assert float((y.double() - ref).abs().max()) < 2e-5, (K, N, act)


@pytest.mark.parametrize('M', [1024, 8192, 32768])      # plain / pre-split one-tile / persistent + A-stationary kernels
@pytest.mark.parametrize('K,N', [(128, 512), (256, 384), (96, 128)])
def test_tc2_engine_three_products(M, K, N):
    """Engine 'tc2' (bf16 x 2 split, three products per MAC): error <= 2e-5 of sum_k |a||b| on forward / dgrad / wgrad
    (the dropped m.m product and the two operand residuals are each <= 2^-18), i.e. far inside the 1e-4 of the TF32-class
    mode the north star allows, and at least 20 x tighter than one bf16 product."""
    rt = _rt()
    g = torch.Generator(device='cuda').manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g, device='cuda')
    w = torch.randn(K, N, generator=g, device='cuda') * 0.1
    dy = torch.randn(M, N, generator=g, device='cuda')
    ref = (x.double() @ w.double(), dy.double() @ w.double().t(), x.double().t() @ dy.double())
    scl = (x.double().abs() @ w.double().abs(), dy.double().abs() @ w.double().abs().t(), x.double().abs().t() @ dy.double().abs())
    err = {}
    for eng in ('tc2', 'bf16'):
        with rt.f32_engine(eng):
            y = _gemm(x, K, 1, w, N, 1, M, N, K)
            dx = _gemm(dy, N, 1, w, 1, N, M, K, N)
            dw = _gemm(x, 1, K, dy, N, 1, K, N, M, splitk=8)
            y2, _ = rt.linear(x, w, None, 'relu')
        err[eng] = [float(((o.double() - r).abs() / s).max()) for o, r, s in zip((y, dx, dw), ref, scl)]
        if eng == 'tc2':
            assert float(((y2.double() - torch.relu(ref[0])).abs() / scl[0]).max()) < 2e-5
    for i, name in enumerate(('forward', 'dgrad', 'wgrad')):
        assert err['tc2'][i] < 2e-5, '%s: %.3g' % (name, err['tc2'][i])
        # (a wgrad whose output is smaller than one 128 x 64 tile runs on the CUDA cores under every engine)
        assert err['bf16'][i] < 1e-6 or err['tc2'][i] * 20 < err['bf16'][i], '%s: tc2 %.3g vs bf16 %.3g' % (name, err['tc2'][i], err['bf16'][i])
    assert rt.get_f32_engine() == 'tc'


@pytest.mark.parametrize('a_unit_k,b_unit_k', [(True, False), (True, True), (False, False), (False, True)])
@pytest.mark.parametrize('M,N,K', [(256, 256, 64), (300, 515, 70), (1000, 333, 200), (4500, 512, 256), (512, 1024, 4100)])
def test_tc2_pair_kernel_layouts(M, N, K, a_unit_k, b_unit_k):
    """The CTA-pair kernel (256 x 256 tiles, cta_group::2) behind the tc2 / bf16 engines, every operand layout, ragged M / N / K,
    split-K (K > 2048 is chunked): error <= 3e-5 of sum_k |a||b| (two-piece operands + one truncating accumulator)."""
    rt = _rt()
    a, b, A, sam, sak, Bm, sbk, sbn = _operands(M, N, K, a_unit_k, b_unit_k, seed=M + N + K)
    bias = torch.randn(N, device='cuda')
    with rt.f32_engine('tc2'):
        C = _gemm(A, sam, sak, Bm, sbk, sbn, M, N, K)
        Cb = _gemm(A, sam, sak, Bm, sbk, sbn, M, N, K, bias=bias)
        Cs = _gemm(A, sam, sak, Bm, sbk, sbn, M, N, K, splitk=3)
    torch.cuda.synchronize()
    ref = a.double() @ b.double()
    scale = a.double().abs() @ b.double().abs()
    for name, out in (('plain', C), ('bias', Cb - bias), ('splitk', Cs)):
        err = float(((out.double() - ref).abs() / scale).max())
        assert err < 3e-5, (name, err)


@pytest.mark.parametrize('act', [None, 'relu', 'tanh'])
def test_linear_epilogues_tc2_pair_kernel(act):
    """t3d_linear_f32 on the CTA-pair kernel (engine tc2, N >= 256, grid >= one CTA per SM): bias, per-group bias, row mask,
    activation, group max (aligned and ragged groups), ragged N -- against float64."""
    rt = _rt()
    G, R, K, N = 80, 256, 192, 333                     # 80 row tiles of 256 x 2 column tiles = 320 CTAs
    M = G * R
    g = torch.Generator(device='cuda').manual_seed(13)
    x = torch.randn(M, K, generator=g, device='cuda')
    w = torch.randn(K, N, generator=g, device='cuda') * 0.1
    b = torch.randn(N, generator=g, device='cuda')
    gb = torch.randn(G, N, generator=g, device='cuda')
    rm = (torch.rand(M, generator=g, device='cuda') < 0.5).float()
    pre = (x.double() @ w.double() + b.double()).view(G, R, N) + gb.double()[:, None, :]
    f = {None: lambda v: v, 'relu': torch.relu, 'tanh': torch.tanh}[act]
    ref = (f(pre) * rm.double().view(G, R, 1)).view(M, N)
    scale = (x.double().abs() @ w.double().abs()).max()
    with rt.f32_engine('tc2'):
        y, _ = rt.linear(x, w, b, act, gbias=gb, rows_per_group=R, rowmask=rm)
        assert float((y.double() - ref).abs().max()) < 3e-5 * float(scale)
        if act == 'relu':
            y2, gmax = rt.linear(x, w, b, act, gbias=gb, rows_per_group=R, rowmask=rm, gmax_groups=G)
            _, gmax_ragged = rt.linear(x[:64 * 300], w, b, act, rows_per_group=300, gmax_groups=64, want_y=False)
            assert torch.equal(y2, y)
            assert torch.equal(gmax, y.view(G, R, N).max(dim=1).values.clamp_min(0.0))
            ref_r = torch.relu(x[:64 * 300].double() @ w.double() + b.double()).view(64, 300, N).max(dim=1).values
            assert float((gmax_ragged.double() - ref_r).abs().max()) < 3e-5 * float(scale)
    assert rt.get_f32_engine() == 'tc'
