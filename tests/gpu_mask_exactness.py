"""Not a test: instance-mask exactness of the B200 path against the oracle (north star: ">= 95 % of frustums bit-exact
on the instance mask").  For each precision mode: fraction of frustums whose 2048 mask bits all agree with the oracle's
fp32 CPU mask, per-point agreement, and the same against the oracle run in float64 (the fp32 oracle's own noise floor).
Weights: synthetic Xavier with the calibrated logit margin (weights.standard_model_F, the bench weights; the flip rate
is invariant to the margin scale k because logits and their rounding noise scale together).  Run under gpurun."""
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


def oracle_logits(variables, pc, dtype):
    vs = VarStore(variables, dtype=dtype)
    with torch.no_grad(), vs.variable_scope('class_agnostic'):
        return osm.v1_inst_seg(torch.as_tensor(pc).to(dtype), None, None, {}, False, vs, scope='inst_seg')


def main():
    B = int(os.environ.get('T3D_MASK_B', '128'))
    dev = 'cuda:0'
    out = []
    for name, margin_std in (('calibrated margin std 2.0 (bench weights)', 2.0),):
        variables, info = weights.standard_model_F(margin_std=margin_std)
        b = synth.make_batch(B, 2048, 6, seed=77)
        ol32 = oracle_logits(variables, b['pc'], torch.float32).numpy()
        m_ref = ol32[:, :, 0] < ol32[:, :, 1]
        ol64 = oracle_logits(variables, b['pc'], torch.float64).numpy()
        a64 = ((ol64[:, :, 0] < ol64[:, :, 1]) == m_ref)
        print(json.dumps(dict(weights=name, mode='oracle fp32 vs oracle fp64 (noise floor of the checker)', frustums=B,
                              frustum_exact=float(a64.all(axis=1).mean()), point_agreement=float(a64.mean()))), flush=True)
        margin = np.abs(ol32[:, :, 1] - ol32[:, :, 0])
        rt.set_default_store(rt.VariableStore(variables, dev))
        pc = torch.as_tensor(b['pc']).to(dev)
        for mode in ('fp32', 'bf16'):
            with rt.precision(mode), torch.no_grad():
                lg = sm.v1_inst_seg(pc, None, None, {}, False, scope='class_agnostic/inst_seg').cpu().numpy()
            m = lg[:, :, 0] < lg[:, :, 1]
            agree = (m == m_ref)
            err = np.abs(lg - ol32).max(axis=2)
            out.append(dict(weights=name, mode=mode, frustums=B, frustum_exact=float(agree.all(axis=1).mean()),
                            point_agreement=float(agree.mean()), flipped_points_per_frustum=float((~agree).sum(axis=1).mean()),
                            masked_in_fraction=float(m_ref.mean()), median_margin=float(np.median(margin)),
                            logit_err_mean=float(err.mean()), logit_err_max=float(err.max()),
                            logit_scale=float(np.abs(ol32).mean())))
            print(json.dumps(out[-1]), flush=True)


if __name__ == '__main__':
    main()
