"""Not a test: forward GEMM of a training layer with / without the lazy-BN extras (t3d_gemm_bn_f32), M = 524288 rows.
Run under gpurun: python tools/gpu_gemm_bn_bench.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transferable3d_b200._lib import ptr, stream, call, gemm_workspace   # noqa: E402
from transferable3d_b200 import runtime as rt   # noqa: E402

rt.set_f32_engine(os.environ.get('ENGINE', 'tc'))

M = 524288
dev = 'cuda:0'
ws = gemm_workspace()
for K, N in ((128, 128), (128, 256), (256, 512), (128, 1024), (512, 256)):
    A = torch.randn(M, K, device=dev)
    W = torch.randn(K, N, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    sc, sh = torch.rand(K, device=dev) + 0.5, torch.randn(K, device=dev)
    C = torch.empty(M, N, device=dev)
    s0, s1, y0 = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.zeros(N, device=dev)
    res = {}
    for name, lazy, stats in (('plain', 0, 0), ('lazyA', 1, 0), ('stats', 0, 1), ('both', 1, 1)):
        def run():
            call('t3d_gemm_bn_f32', ptr(A), K, 1, ptr(sc) if lazy else None, ptr(sh) if lazy else None, ptr(W), N, 1, ptr(C), N, M, N, K, 1,
                 ptr(b), ptr(s0) if stats else None, ptr(s1) if stats else None, ptr(y0) if stats else None, ptr(ws), ws.numel(), stream())
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[name] = round(ms * 1e3)
    res['tflops_plain'] = round(2.0 * M * N * K / (res['plain'] * 1e-6) / 1e12, 1)
    print(json.dumps({'K': K, 'N': N, 'us': res}))
