"""Not a test: stress of the CTA-pair GEMM kernel (tc2 engine): many launches of forward / dgrad / wgrad shaped problems, results
checked against the first launch; run under `timeout` -- a hang shows as a killed process.  Run under gpurun."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transferable3d_b200 import runtime as rt  # noqa: E402
from transferable3d_b200.train_layers import gemm  # noqa: E402

M = int(os.environ.get('M', 65536))
K, N = 256, 512
x = torch.randn(M, K, device='cuda')
w = torch.randn(K, N, device='cuda') * 0.1
dy = torch.randn(M, N, device='cuda')
n = int(os.environ.get('REPS', 1500))
with rt.f32_engine('tc2'):
    ref = (gemm(x, K, 1, w, N, 1, M, N, K), gemm(dy, N, 1, w, 1, N, M, K, N), gemm(x, 1, K, dy, N, 1, K, N, M, splitk=1))
    torch.cuda.synchronize()
    t0 = time.time()
    bad = 0
    for i in range(n):
        out = (gemm(x, K, 1, w, N, 1, M, N, K), gemm(dy, N, 1, w, 1, N, M, K, N), gemm(x, 1, K, dy, N, 1, K, N, M, splitk=1))
        if i % 50 == 0:
            torch.cuda.synchronize()
            bad += int(not torch.equal(out[0], ref[0])) + int(not torch.equal(out[1], ref[1]))
            bad += int(float((out[2] - ref[2]).abs().max()) > 1e-3 * float(ref[2].abs().max()))
            print('iter', i, 'mismatches so far', bad, 'elapsed %.1f s' % (time.time() - t0), flush=True)
    torch.cuda.synchronize()
print('done: %d launches x 3, mismatches %d' % (n, bad))
