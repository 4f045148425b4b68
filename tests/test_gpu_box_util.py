"""-m gpu: oriented 3D IoU, get_3d_box, compute_box3d_iou (get_iou_summary) and the IoU-band perturbation sampler on the
GPU (through the C ABI) against the oracle restatement (oracle/box_util.py, oracle/box_pc_fit_dataset.py).
Tolerance: the GPU computes in fp32, the oracle in float64 -> IoUs within 2e-5 absolute (fp32 round-off of the clip
intersections); index-like outputs (attempt counts) identical except where an IoU falls within that distance of a band edge."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from transferable3d_b200 import box_util as gbu, semisup_v1_sunrgbd as M
    from transferable3d_b200.constants import MEAN_DIMS_ARR

DEV = 'cuda:0'
TOL = 2e-5


def _boxes(B, seed, near=True):
    rng = np.random.RandomState(seed)
    c1 = rng.uniform(-1, 1, (B, 3)) + np.array([0, 0, 3.0])
    s1 = rng.uniform(0.4, 2.5, (B, 3))
    h1 = rng.uniform(-np.pi, np.pi, B)
    d = 0.4 if near else 3.0
    c2 = c1 + rng.uniform(-d, d, (B, 3))
    s2 = s1 * rng.uniform(0.7, 1.4, (B, 3))
    h2 = h1 + rng.uniform(-1.0, 1.0, B)
    return c1, s1, h1, c2, s2, h2


def test_get_3d_box_and_box3d_iou_vs_oracle(built_lib):
    from oracle import box_util as obu
    B = 4096
    c1, s1, h1, c2, s2, h2 = _boxes(B, 1)
    # edge cases: identical, disjoint, touching faces, contained, 90-degree rotation
    c2[0], s2[0], h2[0] = c1[0], s1[0], h1[0]
    c2[1] = c1[1] + 50.0
    c1[2], s1[2], h1[2], c2[2], s2[2], h2[2] = [0, 0, 0], [2, 1, 1], 0.0, [2, 0, 0], [2, 1, 1], 0.0
    c2[3], s2[3], h2[3] = c1[3], s1[3] * 0.25, h1[3]
    c2[4], s2[4], h2[4] = c1[4], s1[4], h1[4] + np.pi / 2
    k1 = gbu.get_3d_box(s1, h1, c1)
    k2 = gbu.get_3d_box(s2, h2, c2)
    i3, i2 = gbu.box3d_iou(k1, k2)
    j3, _ = gbu.get_box3d_iou(c1, s1, h1, c2, s2, h2)
    torch.cuda.synchronize()
    o1 = np.stack([obu.get_3d_box(s1[i], h1[i], c1[i]) for i in range(B)])
    assert np.abs(k1.cpu().numpy() - o1).max() < 1e-5
    ref = np.array([obu.box3d_iou(o1[i], obu.get_3d_box(s2[i], h2[i], c2[i])) for i in range(B)])
    e3, e2 = np.abs(i3.cpu().numpy() - ref[:, 0]), np.abs(i2.cpu().numpy() - ref[:, 1])
    assert e3[1:].max() < TOL and e2[1:].max() < TOL, (e3.argmax(), e3.max(), e2.max())
    assert torch.equal(i3, j3)
    # identical boxes (every edge coincident: the degenerate configuration of the clip) -> 1 up to fp32 round-off
    assert abs(float(i3[0]) - 1.0) < 1e-4 and abs(float(i2[0]) - 1.0) < 1e-4
    assert float(i3[1]) == 0.0 and float(i3[2]) == 0.0
    assert float(i3.max()) <= 1.0 + 1e-6 and float(i3.min()) >= 0.0
    assert (ref[:, 0] > 0.05).mean() > 0.5                       # the comparison is not vacuous


def test_compute_box3d_iou_and_get_iou_summary_vs_oracle(built_lib):
    from oracle import box_util as obu
    B, NH, NS = 512, 12, 10
    rng = np.random.RandomState(2)
    cp, cl = rng.uniform(-1, 1, (B, 3)).astype(np.float32), rng.uniform(-1, 1, (B, 3)).astype(np.float32)
    hl, hr = rng.randn(B, NH).astype(np.float32), (rng.uniform(-0.26, 0.26, (B, NH))).astype(np.float32)
    sl, sr = rng.randn(B, NS).astype(np.float32), (rng.uniform(-0.2, 0.2, (B, NS, 3))).astype(np.float32)
    hl[0, 3] = hl[0, 7] = hl[0].max() + 1.0                      # argmax tie -> first index
    hcl, scl = rng.randint(0, NH, B).astype(np.int32), rng.randint(0, NS, B).astype(np.int32)
    hrl, srl = rng.uniform(-0.26, 0.26, B).astype(np.float32), rng.uniform(-0.2, 0.2, (B, 3)).astype(np.float32)
    o2, o3 = obu.compute_box3d_iou(cp, hl, hr, sl, sr, cl, hcl, hrl, scl, srl)
    D = lambda a: torch.as_tensor(a).to(DEV)
    ep = {}
    g2, g3 = M.get_iou_summary((D(cp), D(sl), D(sr), D(hl), D(hr)), (D(cl), D(scl), D(srl), D(hcl), D(hrl)), ep, name_prefix='W_')
    torch.cuda.synchronize()
    assert np.abs(g2.cpu().numpy() - o2).max() < TOL and np.abs(g3.cpu().numpy() - o3).max() < TOL
    assert ep['W_iou2ds'] is g2 and ep['W_iou3ds'] is g3


def test_perturb_box_to_diff_ious_vs_oracle(built_lib):
    """The GPU sampler reproduces the oracle's philox stream: same attempt count and the same accepted perturbation per
    box (up to fp32 round-off), every accepted IoU inside its band."""
    from oracle import box_pc_fit_dataset as od, box_util as obu
    B = 512
    rng = np.random.RandomState(3)
    c = (rng.uniform(-1, 1, (B, 3)) + np.array([0, 0, 3.0])).astype(np.float32)
    s = (MEAN_DIMS_ARR[rng.randint(0, 10, B)] * rng.uniform(0.8, 1.2, (B, 3))).astype(np.float32)
    h = rng.uniform(-np.pi, np.pi, B).astype(np.float32)
    fit = rng.rand(B) < 0.5
    bounds = np.where(fit[:, None], np.array([[0.7, 1.0]]), np.array([[0.01, 0.25]])).astype(np.float32)
    nc, ns, nh, iou, dc, ds, da, att = gbu.perturb_box_to_diff_ious(c, s, h, bounds, 0.8, 0.2, np.pi, seed=77)
    torch.cuda.synchronize()
    att, iou = att.cpu().numpy(), iou.cpu().numpy()
    assert (att >= 1).all()
    same = 0
    for b in range(B):
        r = od.perturb_box_to_diff_ious(c[b], s[b], float(h[b]), bounds[b].astype(np.float64), 0.8, 0.2, np.pi, rng_mode='philox', seed=77, box_index=b)
        # the accepted box, re-scored by the oracle in float64, lies in the band (up to fp32 round-off at the edges)
        chk, _ = obu.get_box3d_iou(c[b], s[b], float(h[b]), nc[b].cpu().numpy(), ns[b].cpu().numpy(), float(nh[b]))
        assert bounds[b, 0] - TOL < chk < bounds[b, 1] + TOL and abs(chk - iou[b]) < TOL
        if r[7] == att[b]:
            same += 1
            assert np.abs(dc[b].cpu().numpy() - r[4]).max() < 1e-5 and np.abs(ds[b].cpu().numpy() - r[5]).max() < 1e-5
            assert abs(float(da[b]) - r[6]) < 1e-5 and np.abs(nc[b].cpu().numpy() - r[0]).max() < 1e-5
    assert same >= B - 2, same                                   # a band-edge tie within fp32 round-off may shift an attempt
    # an empty band (strict inequalities, lo == hi) gives up after max_attempts instead of spinning for ever
    r = gbu.perturb_box_to_diff_ious(c[:4], s[:4], h[:4], np.array([0.6, 0.6], dtype=np.float32), 0.8, 0.2, np.pi, seed=1, max_attempts=50)
    assert (r[7].cpu().numpy() == -1).all()
