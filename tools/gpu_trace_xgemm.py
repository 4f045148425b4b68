"""Not a test: clock64 timeline of the middle CTA of the xgemm kernel (t3d_set_trace_buffer).  Run under gpurun."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transferable3d_b200 import runtime as rt, _lib  # noqa: E402
from transferable3d_b200.train_layers import gemm  # noqa: E402

SLOTS = 8192


def dump(buf, title):
    b = buf.cpu().numpy().reshape(4, SLOTS)
    print('=== %s' % title)
    t0 = None
    for role, name in enumerate(('loaderA', 'loaderB/mma(pp)', 'mma/epi(pp)')):
        ev = b[role]
        ev = ev[ev != 0]
        if len(ev) == 0:
            continue
        t = (ev >> 8).astype(np.int64)
        tag = (ev & 0xff).astype(np.int64)
        if t0 is None:
            t0 = t[0]
        print('  %-8s ' % name + ' '.join('%02x@%d' % (tg, r) for tg, r in zip(tag[:int(os.environ.get('NEV', 40))], (t - t0)[:int(os.environ.get('NEV', 40))])))


def main():
    M = 524288
    buf = torch.zeros(4 * SLOTS, dtype=torch.int64, device='cuda')
    shapes = [tuple(int(v) for v in kn.split('x')) for kn in os.environ['SHAPES'].split(',')] if os.environ.get('SHAPES') else None
    for K, N in shapes or (((256, 512),) if os.environ.get('ENGINES') else ((128, 128), (256, 512))):
        x = torch.randn(M, K, device='cuda')
        w = torch.randn(K, N, device='cuda') * 0.1
        dy = torch.randn(M, N, device='cuda')
        for eng in (os.environ.get('ENGINES', 'tc,bf16').split(',')):
            with rt.f32_engine(eng):
                gemm(x, K, 1, w, N, 1, M, N, K)
                torch.cuda.synchronize()
                _lib.call('t3d_set_trace_buffer', _lib.ptr(buf))
                buf.zero_()
                gemm(x, K, 1, w, N, 1, M, N, K)
                torch.cuda.synchronize()
                dump(buf, 'fwd %d->%d engine %s (loader: 02 start, 10 loads issued, 20 slot free, 30 stored+arrived, 40 acc full; mma: 10 stage full, 20 issued)' % (K, N, eng))
                buf.zero_()
                gemm(dy, N, 1, w, 1, N, M, K, N)
                torch.cuda.synchronize()
                dump(buf, 'dgrad %d<-%d engine %s' % (K, N, eng))
                buf.zero_()
                gemm(x, 1, K, dy, N, 1, K, N, M, splitk=1)
                torch.cuda.synchronize()
                dump(buf, 'wgrad [%d x %d] engine %s' % (K, N, eng))
                _lib.call('t3d_set_trace_buffer', None)


if __name__ == '__main__':
    main()
