"""Not a test: timings of the widened-row kernels at the cfg3 batch (8192 frustums x 2048 points):
inference scores, prediction -> label, surface loss forward + backward.  Algorithmic bytes per launch / time = GB/s against
the measured HBM peak.  Run under gpurun."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transferable3d_b200 import test_semisup as ts, weak_losses as W, roi_seg_box3d_dataset as ds  # noqa: E402


def timed(fn, n=10):
    for _ in range(3):
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
    B, N = 8192, 2048
    g = torch.Generator(device='cuda').manual_seed(0)
    R = lambda *s: torch.randn(*s, generator=g, device='cuda')
    logits, hs, hr, ss, sr, fp = R(B, N, 2), R(B, 12), R(B, 12), R(B, 10), R(B, 10, 3), torch.rand(B, generator=g, device='cuda')
    ms = timed(lambda: ts.inference_scores(logits, hs, hr, ss, sr, fp))
    byt = B * N * (8 + 1)
    print(json.dumps({'kernel': 'inference_scores_kernel', 'frustums': B, 'ms': round(ms, 4), 'algorithmic_bytes': byt,
                      'GBs': round(byt / ms / 1e6, 1), 'frustums_per_s': round(B / ms * 1e3)}))
    pc = torch.cat([R(B, N, 3) + torch.tensor([0., 0., 3.], device='cuda'), torch.rand(B, N, 3, generator=g, device='cuda')], 2).contiguous()
    soft = torch.rand(B, N, generator=g, device='cuda')
    box = (R(B, 3) * 0.3 + torch.tensor([0., 0., 3.], device='cuda'), torch.rand(B, 3, generator=g, device='cuda') + 0.5, R(B))
    up = R(B)
    ms = timed(lambda: W.get_surface_loss(box, pc, soft, 0.0, 0.9, 0.8, False, (True, False, True), reduce_loss=False))
    byt = B * N * (12 + 4)
    print(json.dumps({'kernel': 'surface_loss_kernel (forward)', 'frustums': B, 'ms': round(ms, 4), 'algorithmic_bytes': byt,
                      'GBs': round(byt / ms / 1e6, 1), 'note': 'xyz of a 6-channel record: 24 B / point cross the bus for 12 used'}))
    ms = timed(lambda: W.get_surface_loss(box, pc, soft, 0.0, 0.9, 0.8, False, (True, False, True), reduce_loss=False, upstream=up, end_points={}))
    byt = B * N * (12 + 4 + 4)
    print(json.dumps({'kernel': 'surface_loss_kernel (forward + backward)', 'frustums': B, 'ms': round(ms, 4), 'algorithmic_bytes': byt,
                      'GBs': round(byt / ms / 1e6, 1)}))
    c, hc, sc = R(B, 3), torch.randint(0, 12, (B,), device='cuda'), torch.randint(0, 10, (B,), device='cuda')
    ms = timed(lambda: ds.from_prediction_to_label_format_batch(c, hc, hr[:, 0].contiguous(), sc, sr[:, 0].contiguous(), fp))
    print(json.dumps({'kernel': 'prediction_to_label_kernel', 'boxes': B, 'ms': round(ms, 4)}))


if __name__ == '__main__':
    main()
