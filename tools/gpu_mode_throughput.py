"""Not a test: frustums/s of the cfg3 pipeline and of model F (+ BoxPC refine) in every precision mode, inputs resident.
Run under gpurun."""
import os
import sys
import json

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transferable3d_b200 import weights, synth, runtime as rt, config, test_semisup as ts, frustum_pointnets_v1 as fpn, model_util as mu  # noqa: E402


def timed(fn, n=3):
    for _ in range(2):
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
    dev = 'cuda:0'
    B = int(os.environ.get('T3D_TP_B', '1024'))
    b = synth.make_batch(64, 2048, 6, seed=3)
    pc = torch.as_tensor(np.tile(b['pc'], (B // 64, 1, 1))).to(dev)
    oh = torch.as_tensor(np.tile(b['one_hot'], (B // 64, 1))).to(dev)
    FLAGS = config.cfg()
    mu.set_resample_rng('philox', seed=5)
    for name, vars_fn, fn in (('cfg3 pipeline (model A weights)', weights.standard_model_A, lambda: fpn.get_model(pc, oh, False)),
                              ('model F + 1 BoxPC refine', weights.standard_model_F, lambda: ts.build_graph(FLAGS, pc, oh))):
        v, _ = vars_fn()
        rt.set_default_store(rt.VariableStore(v, dev))
        for mode in ('bf16', 'f16x2', 'fp32'):
            with rt.precision(mode), torch.no_grad():
                ms = timed(fn)
            print(json.dumps({'graph': name, 'mode': mode, 'frustums': B, 'ms': ms, 'frustums_per_s': B / ms * 1e3}), flush=True)


if __name__ == '__main__':
    main()
