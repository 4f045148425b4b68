"""Not a test: instance-mask exactness and throughput of every precision mode against the oracle (north star: ">= 95 % of
frustums bit-exact on the instance mask").  For each mode: fraction of frustums whose 2048 mask bits all agree with the
oracle's fp32 CPU mask (literal graph: tiled global feature, the reference's own formulation), per-point agreement, logit
error, and seg-chain frustums/s with inputs resident (CUDA events).  f16x2 is swept over the truncation-correction
constant (t3d_set_x2_debias).  The fp32 oracle's own distance to its float64 run on the first frustums is the noise floor
of the checker.  Weights: synthetic Xavier with the calibrated logit margin (weights.standard_model_F, the bench weights;
the flip rate is invariant to the margin scale k because logits and their rounding noise scale together).
Run under gpurun:  T3D_MASK_B=1024 python tools/gpu_mask_exactness.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transferable3d_b200 import weights, synth, runtime as rt, semisup_models as sm   # noqa: E402
from oracle import semisup_models as osm                                              # noqa: E402  (checker)
from oracle.tf_layers import VarStore                                                 # noqa: E402


def oracle_logits(variables, pc, dtype, chunk=32):
    vs = VarStore(variables, dtype=dtype)
    out = []
    with torch.no_grad(), vs.variable_scope('class_agnostic'):
        for i in range(0, pc.shape[0], chunk):
            out.append(osm.v1_inst_seg(torch.as_tensor(pc[i:i + chunk]).to(dtype), None, None, {}, False, vs, scope='inst_seg').float())
    return torch.cat(out).numpy()


def timed(fn, n=5):
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
    B = int(os.environ.get('T3D_MASK_B', '1024'))
    B64 = min(B, int(os.environ.get('T3D_MASK_B64', '128')))
    sweep = [float(x) for x in os.environ.get('T3D_X2_DEBIAS_SWEEP', '0,1.2e-8,2.1e-8,3.0e-8,4.0e-8').split(',')]
    dev = 'cuda:0'
    torch.set_num_threads(os.cpu_count())
    variables, info = weights.standard_model_F(margin_std=2.0)
    b = synth.make_batch(B, 2048, 6, seed=77)
    ol32 = oracle_logits(variables, b['pc'], torch.float32)
    m_ref = ol32[:, :, 0] < ol32[:, :, 1]
    ol64 = oracle_logits(variables, b['pc'][:B64], torch.float64)
    a64 = ((ol64[:, :, 0] < ol64[:, :, 1]) == m_ref[:B64])
    print(json.dumps(dict(mode='oracle fp32 vs oracle fp64 (noise floor of the checker)', frustums=B64,
                          frustum_exact=float(a64.all(axis=1).mean()), point_agreement=float(a64.mean()),
                          logit_err_mean=float(np.abs(ol64 - ol32[:B64]).max(axis=2).mean()))), flush=True)
    margin = np.abs(ol32[:, :, 1] - ol32[:, :, 0])
    rt.set_default_store(rt.VariableStore(variables, dev))
    pc = torch.as_tensor(b['pc']).to(dev)
    default_c = rt.get_x2_debias()
    runs = [('fp32', None), ('bf16', None)] + [('f16x2', c) for c in sweep]
    for mode, c in runs:
        if c is not None:
            rt.set_x2_debias(c)
        with rt.precision(mode), torch.no_grad():
            fn = lambda: sm.v1_inst_seg(pc, None, None, {}, False, scope='class_agnostic/inst_seg')
            lg = fn().cpu().numpy()
            ms = timed(fn)
        m = lg[:, :, 0] < lg[:, :, 1]
        agree = (m == m_ref)
        err = np.abs(lg - ol32).max(axis=2)
        err64 = np.abs(lg[:B64] - ol64).max(axis=2)
        print(json.dumps(dict(mode=mode, x2_debias=c, frustums=B, frustum_exact=float(agree.all(axis=1).mean()),
                              point_agreement=float(agree.mean()), flipped_points_per_frustum=float((~agree).sum(axis=1).mean()),
                              masked_in_fraction=float(m_ref.mean()), median_margin=float(np.median(margin)),
                              logit_err_mean=float(err.mean()), logit_err_max=float(err.max()),
                              logit_err_mean_vs_fp64=float(err64.mean()), logit_scale=float(np.abs(ol32).mean()),
                              seg_ms=ms, seg_frustums_per_s=B / ms * 1e3)), flush=True)
    rt.set_x2_debias(default_c)


if __name__ == '__main__':
    main()
