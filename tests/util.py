"""Shared helpers for the parity tests (tests may import the oracle; the product may not)."""
import numpy as np
import torch

from transferable3d_b200 import weights, synth, config


def model_F_setup(B, N=2048, seed=1234, margin=True):
    if margin:
        variables, info = weights.standard_model_F()
    else:
        variables, info = weights.make_weights_model_F(), {}
    batch = synth.make_batch(B, N, 6, seed=seed)
    return variables, batch, config.cfg(), info


def oracle_model_F(variables, batch, FLAGS, dtype=torch.float32, literal=True):
    from oracle.tf_layers import VarStore
    from oracle import test_semisup
    vs = VarStore(variables, dtype=dtype)
    vs.literal = literal
    with torch.no_grad():
        logits, ep = test_semisup.run_graph(vs, FLAGS, torch.as_tensor(batch['pc']).to(dtype),
                                            torch.as_tensor(batch['one_hot']).to(dtype))
    return logits, ep


def err_stats(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    return dict(max_abs=float(d.max()), mean_abs=float(d.mean()), ref_scale=float(np.abs(b).mean()),
                max_rel=float((d / (np.abs(b) + 1e-3)).max()))


def assert_close(a, b, rel, abs_, what='', frac=1.0):
    """|a-b| <= abs_ + rel*|b| for at least `frac` of the elements."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    ok = np.abs(a - b) <= abs_ + rel * np.abs(b)
    got = ok.mean()
    assert got >= frac, '%s: only %.5f of elements within rel=%g abs=%g (%s)' % (what, got, rel, abs_, err_stats(a, b))
