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


def oracle_model_F(variables, batch, FLAGS, dtype=torch.float32, literal=True, oracle_mask=None):
    """Oracle run of the test graph.  oracle_mask (B,N) 0/1: continue the oracle from a GIVEN mask (the reference's own
    oracle_mask input, semisup_v1_sunrgbd.py:161-162) -- used to compare everything downstream of the strict logit
    compare on identical masks; the returned logits are then the stacked mask, the seg logits stay in
    ep['_oracle_seg_logits'] when computed separately."""
    from oracle.tf_layers import VarStore
    from oracle import test_semisup
    vs = VarStore(variables, dtype=dtype)
    vs.literal = literal
    om = None if oracle_mask is None else torch.as_tensor(np.asarray(oracle_mask)).to(dtype)
    with torch.no_grad():
        logits, ep = test_semisup.run_graph(vs, FLAGS, torch.as_tensor(batch['pc']).to(dtype),
                                            torch.as_tensor(batch['one_hot']).to(dtype), oracle_mask=om)
    return logits, ep


def oracle_seg_logits(variables, pc, one_hot=None, scope='class_agnostic/inst_seg', chunk=64, dtype=torch.float32):
    """Oracle mask logits of many frustums in chunks (folded conv6: algebraically identical, a third of the memory)."""
    from oracle.tf_layers import VarStore
    from oracle import semisup_models as osm
    vs = VarStore(variables, dtype=dtype)
    vs.literal = False
    out = []
    with torch.no_grad():
        for i in range(0, pc.shape[0], chunk):
            oh = None if one_hot is None else torch.as_tensor(one_hot[i:i + chunk]).to(dtype)
            parts = scope.split('/')
            ctx = [vs.variable_scope(p) for p in parts[:-1]]
            for c in ctx:
                c.__enter__()
            try:
                out.append(osm.v1_inst_seg(torch.as_tensor(pc[i:i + chunk]).to(dtype), None, oh, {}, False, vs, scope=parts[-1]))
            finally:
                for c in reversed(ctx):
                    c.__exit__(None, None, None)
    return torch.cat(out)


def oracle_cfg3_from_logits(variables, pc, one_hot, logits, seed, dtype=torch.float32):
    """F-PointNet v1 pipeline downstream of GIVEN mask logits (oracle/frustum_pointnets_v1.py): the oracle continued from
    the GPU's own logits, so that masks and resampled indices are identical and every later output is comparable."""
    from oracle.tf_layers import VarStore
    from oracle import frustum_pointnets_v1 as ofpn
    vs = VarStore(variables, dtype=dtype)
    with torch.no_grad():
        return ofpn.get_model(vs, torch.as_tensor(pc).to(dtype), torch.as_tensor(one_hot).to(dtype), seed=seed,
                              logits=torch.as_tensor(logits).to(dtype))


def frac_within(a, b, rel, abs_):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float((np.abs(a - b) <= abs_ + rel * np.abs(b)).mean())


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


def boxpc_seed_without_pool_ties(make_setup, B, N, conv_last, first_seed=3, min_gap=6e-6):
    """The gradient of a max-pool is discontinuous where two rows tie: when the two largest values of a pooled column agree
    to ~1e-6 (fp32 round-off), the arg-max -- and with it a visible part of the upstream gradients -- depends on the
    summation order of the batch statistics, on the GPU (atomics) and in the fp32 oracle alike.  Parity of gradients is
    only defined away from such ties, so the training-step tests scan seeds until the pooled layer of the GPU forward pass
    has no top-2 gap below `min_gap` (relative; the measured run-to-run movement of a gap is ~1e-6, and with B x C pooled
    columns a few seeds in ten have no gap below 6e-6).  make_setup(seed) -> (variables, feed, masks, FLAGS)."""
    from transferable3d_b200 import train_boxpc as tb
    from transferable3d_b200.train_layers import Lazy
    for seed in range(first_seed, first_seed + 80):
        v, feed, masks, FLAGS = make_setup(seed)
        g = tb.BoxPCTrainGraph(v, FLAGS, B, N, 6, 'cuda:0')
        g.forward_backward(feed, masks)
        layer = g.layers[conv_last]
        C = layer.N
        x = Lazy(layer).materialize().reshape(B, N, C)
        top2 = torch.topk(x, 2, dim=1).values
        rel = torch.where(top2[:, 0] > 0, (top2[:, 0] - top2[:, 1]) / top2[:, 0].clamp_min(1e-30), torch.ones_like(top2[:, 0]))
        if float(rel.min()) >= min_gap:
            return seed
    raise AssertionError('no seed without a pooled near-tie found')


def assert_grad_close(name, got, ref32, ref64, scale_floor=0.0):
    """Gradient of one variable against the oracle.  Element-wise: mean error within 5x the fp32 oracle's own distance to its
    float64 run + 2e-3 of the scale, max error within 5e-2 of the scale.  The gradient of a ReLU / max-pool network is
    discontinuous where a pre-activation is within round-off of zero: a handful of the B*N x C elements of every layer flip
    from run to run with the summation order of the batch statistics (atomics), each flip moving a column of the layer's
    weight gradient and, slightly, everything below it.  When the element-wise bound is exceeded the tensor must still
    agree as a whole: relative Frobenius error <= 1e-2 and cosine >= 0.9999 (a wrong kernel is off by far more; a tie in
    the max-pool, which moves gradients by percents, is excluded by boxpc_seed_without_pool_ties)."""
    got, ref32, ref64 = (np.asarray(a, dtype=np.float64).reshape(-1) for a in (got, ref32, ref64))
    s = err_stats(got, ref64)
    floor = err_stats(ref32, ref64)['mean_abs']
    scale = max(s['ref_scale'], scale_floor, 1e-7)
    assert np.isfinite(got).all(), name
    if s['mean_abs'] <= 5 * floor + 2e-3 * scale + 1e-8 and s['max_abs'] <= 5e-2 * scale + 1e-7:
        return
    nr = float(np.linalg.norm(ref64))
    assert nr > 1e-9, (name, s, floor)
    rel = float(np.linalg.norm(got - ref64)) / nr
    cos = float(np.dot(got, ref64)) / (float(np.linalg.norm(got)) * nr + 1e-300)
    assert rel <= 1e-2 and cos >= 0.9999, (name, s, floor, rel, cos)
