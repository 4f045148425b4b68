"""Generates tests/golden/model_F_tiny.npz from the oracle: a regression fixture of the ORACLE itself.  The fixtures that
pin the oracle against the reference's source are the ref_*.npz files (make_reference_golden.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.tf_layers import VarStore  # noqa: E402
from oracle import test_semisup  # noqa: E402
from transferable3d_b200 import weights, synth, config  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    B, N, wseed, dseed = 2, 256, 11, 2024
    variables = weights.make_weights_model_F(seed=wseed)
    b = synth.make_batch(B, N, 6, seed=dseed)
    vs = VarStore(variables)
    with torch.no_grad():
        logits, ep = test_semisup.run_graph(vs, config.cfg(), torch.as_tensor(b['pc']), torch.as_tensor(b['one_hot']))
    out = dict(B=B, N=N, weight_seed=wseed, data_seed=dseed, logits=logits.numpy())
    for k in ('F2_center', 'F_size_residuals', 'F_heading_scores', 'boxpc_fit_prob', 'stage1_center'):
        out[k] = ep[k].numpy()
    np.savez_compressed(os.path.join(HERE, 'model_F_tiny.npz'), **out)
    print('wrote model_F_tiny.npz')


if __name__ == '__main__':
    main()
