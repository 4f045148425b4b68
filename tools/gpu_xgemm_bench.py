"""Not a test: fp32 GEMM throughput of the two engines (tcgen05 bf16 x 3 vs CUDA-core SGEMM) on the layer shapes of the
training steps (cfg4 / cfg5: 524288 rows) -- forward, dgrad, wgrad.  Run under gpurun."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transferable3d_b200 import runtime as rt  # noqa: E402
from transferable3d_b200.train_layers import gemm, splitk_for  # noqa: E402


def timed(fn, n=int(os.environ.get('ITERS', 5))):
    for _ in range(int(os.environ.get('WARM', 2))):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    M = int(os.environ.get('ROWS', 256 * 2048))
    for K, N in ((128, 128), (128, 256), (256, 512), (64, 512), (128, 1024)):
        x = torch.randn(M, K, device='cuda')
        w = torch.randn(K, N, device='cuda') * 0.1
        dy = torch.randn(M, N, device='cuda')
        y = torch.empty(M, N, device='cuda')
        dx = torch.empty(M, K, device='cuda')
        dw = torch.empty(K, N, device='cuda')
        sk = splitk_for(K, N, M)
        flops = 2.0 * M * K * N
        for eng in os.environ.get('ENGINES', 'tc,simt,bf16').split(','):
            with rt.f32_engine(eng):
                t_f = timed(lambda: gemm(x, K, 1, w, N, 1, M, N, K, out=y))
                t_d = timed(lambda: gemm(dy, N, 1, w, 1, N, M, K, N, out=dx))
                t_w = timed(lambda: gemm(x, 1, K, dy, N, 1, K, N, M, splitk=sk, out=dw))
            print(json.dumps({'K': K, 'N': N, 'rows': M, 'engine': eng, 'splitk': sk,
                              'fwd_ms': round(t_f, 4), 'dgrad_ms': round(t_d, 4), 'wgrad_ms': round(t_w, 4),
                              'fwd_tflops': round(flops / t_f / 1e9, 1), 'dgrad_tflops': round(flops / t_d / 1e9, 1),
                              'wgrad_tflops': round(flops / t_w / 1e9, 1),
                              'fwd_GBs': round(4.0 * M * (K + N) / t_f / 1e6, 0)}), flush=True)


if __name__ == '__main__':
    main()
