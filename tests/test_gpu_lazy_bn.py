"""-m gpu: the lazy batch-norm operators of the training steps (include/t3d_b200.h "lazy batch norm") against float64
restatements of tf_util.conv2d + batch_norm_template in training mode (models/tf_util.py:1258-1323, 1645-1664):

  t3d_gemm_bn_f32   forward with the BN map of the previous layer applied to A in the loader and the column statistics of the
                    output accumulated in the epilogue -- on each kernel it dispatches to (A-stationary, persistent, one-tile,
                    first-layer); wgrad with the map applied per row of X^T
  t3d_row0, t3d_bn_finalize_affine, t3d_colstats_lazy, t3d_bn_backward_lazy, t3d_maxpool_lazy_fwd
  t3d_pool_bn_backward  == t3d_maxpool_masked_bwd + t3d_colstats + t3d_bn_backward of the stored-output path
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from transferable3d_b200 import _lib
    from transferable3d_b200._lib import ptr, stream, call, gemm_workspace

DEV = 'cuda:0'


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(DEV)


# (M, K, N): K <= 128 & >= 4 column tiles -> A-stationary; K <= 128 -> persistent; K > 128 -> one-tile pre-split kernel;
# ragged M (rows past the end must not enter the statistics); K = 12 -> the first-layer kernel (statistics only)
@pytest.mark.parametrize('M,K,N,lazy_a', [(4096 + 77, 128, 1024, True), (8192, 64, 512, True), (4096, 128, 128, True),
                                          (5000, 128, 256, True), (4224, 256, 512, True), (4096, 512, 256, True),
                                          (6000, 64, 64, False), (4100, 12, 128, False), (4096, 6, 64, False)])
def test_gemm_bn_forward(M, K, N, lazy_a, built_lib):
    rng = np.random.RandomState(M + K + N)
    yp = rng.randn(M, K).astype(np.float32) * 2 + 0.5
    sc = (rng.rand(K).astype(np.float32) + 0.5) * np.where(rng.rand(K) < 0.2, -1, 1).astype(np.float32)
    sh = rng.randn(K).astype(np.float32) * 0.5
    W = (rng.randn(K, N) / np.sqrt(K)).astype(np.float32)
    b = (rng.randn(N) * 3).astype(np.float32)          # |mean| >> std for some columns: the shift matters
    kind = _lib.load().t3d_gemm_bn_supported(M, N, K, 0)
    assert kind == (1 if K >= 32 else 2)
    x64 = np.maximum(sc.astype(np.float64) * yp + sh, 0) if lazy_a else yp.astype(np.float64)
    ref = x64 @ W.astype(np.float64) + b
    d_yp, d_sc, d_sh, d_W, d_b = _dev(yp), _dev(sc), _dev(sh), _dev(W), _dev(b)
    y0 = torch.empty(N, device=DEV)
    call('t3d_row0', ptr(d_yp), ptr(d_sc) if lazy_a else None, ptr(d_sh) if lazy_a else None, ptr(d_W), N, ptr(d_b), K, N, ptr(y0),
         stream())
    assert np.allclose(y0.cpu().numpy(), ref[0], rtol=1e-5, atol=1e-4)
    C = torch.full((M, N), float('nan'), device=DEV)
    s0 = torch.full((N,), 7.0, device=DEV)             # zeroed inside
    s1 = torch.full((N,), 7.0, device=DEV)
    ws = gemm_workspace()
    call('t3d_gemm_bn_f32', ptr(d_yp), K, 1, ptr(d_sc) if lazy_a else None, ptr(d_sh) if lazy_a else None, ptr(d_W), N, 1, ptr(C), N,
         M, N, K, 1, ptr(d_b), ptr(s0), ptr(s1), ptr(y0), ptr(ws), ws.numel(), stream())
    got = C.cpu().numpy().astype(np.float64)
    scale = np.abs(ref).mean()
    assert np.isfinite(got).all() and np.abs(got - ref).max() <= 2e-5 * np.sqrt(K) * scale
    d = got - y0.cpu().numpy().astype(np.float64)      # the statistics the epilogue should have produced, from ITS output
    assert np.allclose(s0.cpu().numpy(), d.sum(0), rtol=2e-4, atol=2e-3 * np.sqrt(M))
    assert np.allclose(s1.cpu().numpy(), (d * d).sum(0), rtol=2e-4, atol=1e-2)
    # finalize: mean / rstd / folded map / moving statistics against float64 on the reference output
    gamma, beta = _dev(rng.rand(N) + 0.5), _dev(rng.randn(N))
    mm, mv = _dev(rng.randn(N)), _dev(rng.rand(N) + 0.5)
    mm0, mv0 = mm.cpu().numpy().copy(), mv.cpu().numpy().copy()
    mean, rstd, a_sc, a_sh = (torch.empty(N, device=DEV) for _ in range(4))
    call('t3d_bn_finalize_affine', ptr(s0), ptr(s1), ptr(y0), M, N, 1e-3, 0.9, ptr(gamma), ptr(beta), ptr(mean), ptr(rstd), ptr(a_sc),
         ptr(a_sh), ptr(mm), ptr(mv), stream())
    mu, var = ref.mean(0), ref.var(0)
    assert np.allclose(mean.cpu().numpy(), mu, rtol=1e-5, atol=1e-5 * scale)
    assert np.allclose(rstd.cpu().numpy(), 1 / np.sqrt(var + 1e-3), rtol=2e-4)
    g64, b64 = gamma.cpu().numpy().astype(np.float64), beta.cpu().numpy().astype(np.float64)
    assert np.allclose(a_sc.cpu().numpy(), g64 / np.sqrt(var + 1e-3), rtol=2e-4)
    assert np.allclose(a_sh.cpu().numpy(), b64 - mu * g64 / np.sqrt(var + 1e-3), rtol=1e-3, atol=1e-3)
    assert np.allclose(mm.cpu().numpy(), 0.9 * mm0 + 0.1 * mu, rtol=1e-5, atol=1e-5 * scale)
    assert np.allclose(mv.cpu().numpy(), 0.9 * mv0 + 0.1 * var * M / (M - 1), rtol=2e-4)


@pytest.mark.parametrize('M,K,N', [(8192, 128, 256), (4096 + 64, 256, 512), (16384, 128, 128)])
def test_gemm_bn_wgrad(M, K, N, built_lib):
    """dW[K, N] = relu(sc * Yprev + sh)^T dY, the BN map applied per ROW of the row-contiguous operand X^T."""
    rng = np.random.RandomState(M + K)
    yp = rng.randn(M, K).astype(np.float32)
    sc = (rng.rand(K).astype(np.float32) + 0.5) * np.where(rng.rand(K) < 0.2, -1, 1).astype(np.float32)
    sh = rng.randn(K).astype(np.float32) * 0.5
    dY = rng.randn(M, N).astype(np.float32)
    assert _lib.load().t3d_gemm_bn_supported(K, N, M, 1) == 1
    ref = np.maximum(sc.astype(np.float64) * yp + sh, 0).T @ dY.astype(np.float64)
    dW = torch.empty(K, N, device=DEV)
    ws = gemm_workspace()
    d_yp, d_sc, d_sh, d_dY = _dev(yp), _dev(sc), _dev(sh), _dev(dY)
    for splitk in (1, 8):
        call('t3d_gemm_bn_f32', ptr(d_yp), 1, K, ptr(d_sc), ptr(d_sh), ptr(d_dY), N, 1, ptr(dW), N, K, N, M, splitk, None, None, None, None,
             ptr(ws), ws.numel(), stream())
        assert np.abs(dW.cpu().numpy() - ref).max() <= 2e-5 * np.sqrt(M) * np.abs(ref).mean() / 10 + 1e-3


def test_gemm_bn_rejects_what_it_cannot_serve(built_lib):
    lib = _lib.load()
    assert lib.t3d_gemm_bn_supported(256, 128, 128, 0) == 0            # M < 4096: no pre-split path
    assert lib.t3d_gemm_bn_supported(8192, 128, 48, 0) == 0            # K % 32 != 0
    a = torch.zeros(256, 128, device=DEV)
    w = torch.zeros(128, 128, device=DEV)
    c = torch.zeros(256, 128, device=DEV)
    v = torch.zeros(128, device=DEV)
    ws = gemm_workspace()
    rc = lib.t3d_gemm_bn_f32(ptr(a), 128, 1, ptr(v), ptr(v), ptr(w), 128, 1, ptr(c), 128, 256, 128, 128, 1, None, None, None, None, ptr(ws),
                             ws.numel(), stream())
    assert rc != 0


@pytest.mark.parametrize('B,N,C,masked', [(8, 256, 512, False), (16, 512, 256, True), (4, 300, 64, True)])
def test_pool_bn_backward_and_lazy_ops(B, N, C, masked, built_lib):
    rng = np.random.RandomState(B * N + C)
    M = B * N
    y = _dev(rng.randn(M, C) * 1.5 + 0.3)
    gamma = _dev((rng.rand(C) + 0.5) * np.where(rng.rand(C) < 0.2, -1, 1))
    beta = _dev(rng.randn(C) * 0.5)
    rowmask = _dev(rng.rand(M) < 0.6) if masked else None
    s0, s1, mean, rstd, a_sc, a_sh = (torch.empty(C, device=DEV) for _ in range(6))
    call('t3d_colstats', ptr(y), None, ptr(y), None, None, ptr(s0), ptr(s1), M, C, 0, 0, stream())
    call('t3d_bn_finalize_affine', ptr(s0), ptr(s1), ptr(y), M, C, 1e-3, 0.5, ptr(gamma), ptr(beta), ptr(mean), ptr(rstd), ptr(a_sc),
         ptr(a_sh), None, None, stream())
    out = torch.empty_like(y)
    call('t3d_bn_apply', ptr(y), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), ptr(out), M, C, 1, stream())
    # lazy max-pool == max-pool of the materialised activation
    p1, a1 = torch.empty(B, C, device=DEV), torch.empty(B, C, dtype=torch.int32, device=DEV)
    p2, a2 = torch.empty(B, C, device=DEV), torch.empty(B, C, dtype=torch.int32, device=DEV)
    call('t3d_maxpool_masked_fwd', ptr(out), ptr(rowmask), B, N, C, ptr(p1), ptr(a1), stream())
    call('t3d_maxpool_lazy_fwd', ptr(y), ptr(a_sc), ptr(a_sh), ptr(rowmask), B, N, C, ptr(p2), ptr(a2), stream())
    assert torch.allclose(p1, p2, rtol=1e-5, atol=1e-5)
    assert (a1 == a2).float().mean() > 0.995           # the folded map rounds differently: near-ties may pick another row
    g = _dev(rng.randn(B, C))
    # stored-output path
    dx = torch.empty(M, C, device=DEV)
    call('t3d_maxpool_masked_bwd', ptr(g), ptr(a2), ptr(rowmask), B, N, C, ptr(dx), stream())
    dx_lazy = dx.clone()
    r1, r2 = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    call('t3d_colstats', ptr(dx), ptr(out), ptr(y), ptr(mean), ptr(rstd), ptr(r1), ptr(r2), M, C, 1, 1, stream())
    call('t3d_bn_backward', ptr(dx), ptr(out), ptr(y), ptr(mean), ptr(rstd), ptr(gamma), ptr(r1), ptr(r2), M, C, 1, stream())
    # lazy dense path (mask recomputed from y)
    l1, l2 = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    call('t3d_colstats_lazy', ptr(dx_lazy), ptr(y), ptr(mean), ptr(rstd), ptr(a_sc), ptr(a_sh), ptr(l1), ptr(l2), M, C, stream())
    call('t3d_bn_backward_lazy', ptr(dx_lazy), ptr(y), ptr(mean), ptr(rstd), ptr(gamma), ptr(a_sc), ptr(a_sh), ptr(l1), ptr(l2), M, C,
         stream())
    # pooled path
    q1, q2, dY = torch.empty(C, device=DEV), torch.empty(C, device=DEV), torch.empty(M, C, device=DEV)
    call('t3d_pool_bn_backward', ptr(g), ptr(a2), ptr(rowmask), ptr(y), ptr(mean), ptr(rstd), ptr(gamma), ptr(a_sc), ptr(a_sh), B, N, C,
         ptr(q1), ptr(q2), ptr(dY), stream())
    torch.cuda.synchronize()
    tol = dict(rtol=1e-4, atol=1e-5 * float(g.abs().max()))
    for a, b_ in ((r1, l1), (r2, l2), (r1, q1), (r2, q2)):
        assert torch.allclose(a, b_, rtol=1e-4, atol=1e-4)
    assert torch.allclose(dx, dx_lazy, **tol)
    assert torch.allclose(dx, dY, **tol)


@pytest.mark.parametrize('B,N,C,lazy,masked', [(4, 2048, 512, True, True), (3, 1000, 300, False, False), (2, 4096, 64, False, True),
                                               (8, 256, 1024, True, False)])
def test_maxpool_row_split_equals_serial(B, N, C, lazy, masked, built_lib):
    """t3d_maxpool_fwd_ws (rows split across blocks, 64-bit atomic max on (value, ~row)) == the serial kernel, including
    ties (first row wins), negative values and all-masked groups."""
    rng = np.random.RandomState(N + C)
    x = np.round(rng.randn(B * N, C) * 2) / 2                       # many exact ties
    x[:N, :7] = -3.0                                                # constant negative columns in the first group: row 0 wins
    xd = _dev(x)
    rm = rng.rand(B * N) < 0.5
    rm[N:2 * N] = False                                             # one group fully masked
    rowmask = _dev(rm) if masked else None
    sc = _dev((rng.rand(C) + 0.5) * np.where(rng.rand(C) < 0.3, -1, 1)) if lazy else None
    sh = _dev(rng.randn(C)) if lazy else None
    p1, a1 = torch.empty(B, C, device=DEV), torch.empty(B, C, dtype=torch.int32, device=DEV)
    p2, a2 = torch.empty(B, C, device=DEV), torch.empty(B, C, dtype=torch.int32, device=DEV)
    keys = torch.empty(B, C, dtype=torch.int64, device=DEV)
    call('t3d_maxpool_fwd_ws', ptr(xd), ptr(sc), ptr(sh), ptr(rowmask), B, N, C, ptr(p1), ptr(a1), None, stream())
    call('t3d_maxpool_fwd_ws', ptr(xd), ptr(sc), ptr(sh), ptr(rowmask), B, N, C, ptr(p2), ptr(a2), ptr(keys), stream())
    torch.cuda.synchronize()
    assert torch.equal(p1, p2) and torch.equal(a1, a2)
    # and against numpy
    v = x.reshape(B, N, C).astype(np.float32)
    if lazy:
        v = np.maximum(np.float32(sc.cpu().numpy()) * v + np.float32(sh.cpu().numpy()), 0)
    if masked:
        v = v * rm.reshape(B, N, 1).astype(np.float32)
    assert np.allclose(p2.cpu().numpy(), v.max(1), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('B,N,C', [(5, 2048, 512), (3, 300, 512), (2, 64, 256)])
def test_pool_rows_gather_scatter(B, N, C, built_lib):
    """Row compaction behind a max-pool (t3d_pool_rows / t3d_gather_rows / t3d_scatter_pool_grad): the compacted tensors hold
    exactly what the dense scatter (t3d_maxpool_bwd) and a row gather would."""
    from transferable3d_b200.train_layers import pool_rows, gather_rows, scatter_pool_grad
    rng = np.random.RandomState(B + N)
    arg = rng.randint(0, N, size=(B, C)).astype(np.int32)
    arg[0, :] = 7                                                   # one frustum whose channels all pick the same row
    d_arg = torch.as_tensor(arg).to(DEV)
    rows, slot, count, S = pool_rows(d_arg, B, N, C)
    assert S == min(C, N)
    rows_h, slot_h, count_h = rows.cpu().numpy(), slot.cpu().numpy(), count.cpu().numpy()
    for b in range(B):
        uniq = np.unique(arg[b])
        assert count_h[b] == len(uniq)
        assert np.array_equal(rows_h[b, :len(uniq)], uniq) and (rows_h[b, len(uniq):] == -1).all()
        assert np.array_equal(rows_h[b, slot_h[b]], arg[b])
    src = _dev(rng.randn(B * N, 12))
    got = gather_rows(src, rows, B, N, S).cpu().numpy().reshape(B, S, 12)
    src_h = src.cpu().numpy().reshape(B, N, 12)
    for b in range(B):
        n = count_h[b]
        assert np.array_equal(got[b, :n], src_h[b, rows_h[b, :n]]) and (got[b, n:] == 0).all()
    src6 = _dev(rng.randn(B * N, 6))                                # C % 4 != 0: scalar path
    got6 = gather_rows(src6, rows, B, N, S).cpu().numpy().reshape(B, S, 6)
    assert np.array_equal(got6[1, :count_h[1]], src6.cpu().numpy().reshape(B, N, 6)[1, rows_h[1, :count_h[1]]])
    g = _dev(rng.randn(B, C))
    comp = scatter_pool_grad(g, slot, B, C, S).cpu().numpy().reshape(B, S, C)
    dense = torch.empty(B * N, C, device=DEV)
    call('t3d_maxpool_bwd', ptr(g), ptr(d_arg), B, N, C, ptr(dense), stream())
    dense = dense.cpu().numpy().reshape(B, N, C)
    for b in range(B):
        n = count_h[b]
        assert np.array_equal(comp[b, :n], dense[b, rows_h[b, :n]]) and (comp[b, n:] == 0).all()
        rest = np.ones(N, bool)
        rest[rows_h[b, :n]] = False
        assert (dense[b, rest] == 0).all()


@pytest.mark.parametrize('B,N,K,C', [(6, 1024, 128, 1024), (40, 128, 64, 128), (5, 2048, 256, 512)])
def test_gemm_bn_pool_forward_only_layer(B, N, K, C, built_lib):
    """t3d_gemm_bn_pool_f32 + t3d_pool_bn_finish == GEMM -> batch statistics -> BN -> ReLU -> max over the N rows of each group,
    without the B*N x C output; channels with a negative BN scale take the group minimum."""
    rng = np.random.RandomState(B + K + C)
    M = B * N
    yp = rng.randn(M, K).astype(np.float32)
    sc = ((rng.rand(K) + 0.5) * np.where(rng.rand(K) < 0.2, -1, 1)).astype(np.float32)
    sh = (rng.randn(K) * 0.5).astype(np.float32)
    W = (rng.randn(K, C) / np.sqrt(K)).astype(np.float32)
    b = rng.randn(C).astype(np.float32)
    gamma = ((rng.rand(C) + 0.5) * np.where(rng.rand(C) < 0.3, -1, 1)).astype(np.float32)
    beta = (rng.randn(C) * 0.3).astype(np.float32)
    x64 = np.maximum(sc.astype(np.float64) * yp + sh, 0)
    y = x64 @ W.astype(np.float64) + b
    mu, var = y.mean(0), y.var(0)
    out = np.maximum(gamma * (y - mu) / np.sqrt(var + 1e-3) + beta, 0).reshape(B, N, C).max(1)
    d = [_dev(a) for a in (yp, sc, sh, W, b, gamma, beta)]
    E = lambda: torch.empty(C, device=DEV)
    y0, s0, s1, mean, rstd, a_sc, a_sh = E(), E(), E(), E(), E(), E(), E()
    call('t3d_row0', ptr(d[0]), ptr(d[1]), ptr(d[2]), ptr(d[3]), C, ptr(d[4]), K, C, ptr(y0), stream())
    kmax = torch.empty((B, C), dtype=torch.int32, device=DEV)
    kmin = torch.empty((B, C), dtype=torch.int32, device=DEV)
    ws = gemm_workspace()
    call('t3d_gemm_bn_pool_f32', ptr(d[0]), K, ptr(d[1]), ptr(d[2]), ptr(d[3]), C, M, C, K, ptr(d[4]), ptr(s0), ptr(s1), ptr(y0), N,
         ptr(kmax), ptr(kmin), ptr(ws), ws.numel(), stream())
    call('t3d_bn_finalize_affine', ptr(s0), ptr(s1), ptr(y0), M, C, 1e-3, 0.5, ptr(d[5]), ptr(d[6]), ptr(mean), ptr(rstd), ptr(a_sc),
         ptr(a_sh), None, None, stream())
    pooled = torch.empty((B, C), device=DEV)
    call('t3d_pool_bn_finish', ptr(kmax), ptr(kmin), ptr(a_sc), ptr(a_sh), B, C, ptr(pooled), stream())
    got = pooled.cpu().numpy()
    assert np.allclose(mean.cpu().numpy(), mu, rtol=1e-5, atol=1e-5)
    assert np.abs(got - out).max() <= 2e-4 * max(1.0, np.abs(out).max()), np.abs(got - out).max()
