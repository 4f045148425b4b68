"""-m gpu: weak_losses.get_surface_loss (first item) through the C ABI against the oracle restatement
(models/weak_losses.py:240-265, models/tf_util.py:610-720): forward values, and the one-pass gradients against the
oracle's autograd for every train_box flag combination."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(B, N, seed):
    rng = np.random.RandomState(seed)
    center = rng.randn(B, 3).astype(np.float32) * 0.3 + np.array([0, 0, 3], np.float32)
    dims = rng.uniform(0.4, 2.0, (B, 3)).astype(np.float32)
    orient = rng.uniform(-np.pi, np.pi, B).astype(np.float32)
    pc = (center[:, None, :] + rng.randn(B, N, 3) * rng.uniform(0.2, 1.5, (B, 1, 1))).astype(np.float32)
    pc = np.concatenate([pc, rng.rand(B, N, 3).astype(np.float32)], axis=2)           # xyz + rgb
    soft = rng.rand(B, N).astype(np.float32)
    up = rng.randn(B).astype(np.float32)
    return center, dims, orient, pc, soft, up


@pytest.mark.parametrize('train_box', [(True, False, True), (True, True, True), (False, True, False)])
@pytest.mark.parametrize('margin,scale', [(0.0, 0.9), (0.05, 1.0)])
def test_surface_loss_forward_and_gradients(train_box, margin, scale):
    from transferable3d_b200 import weak_losses as W
    from oracle import weak_losses as OW
    B, N = 6, 1024
    center, dims, orient, pc, soft, up = _case(B, N, seed=int(margin * 100) + sum(train_box))
    T = lambda a: torch.as_tensor(a, dtype=torch.float64)
    oc, od, oo, osm = T(center).requires_grad_(), T(dims).requires_grad_(), T(orient).requires_grad_(), T(soft).requires_grad_()
    oloss = OW.get_surface_loss((oc, od, oo), T(pc[:, :, :3]), osm, margin, scale, 0.8, False, train_box, reduce_loss=False)
    (oloss * T(up)).sum().backward()
    D = lambda a: torch.as_tensor(a).cuda()
    ep = {}
    loss = W.get_surface_loss((D(center), D(dims), D(orient)), D(pc), D(soft), margin, scale, 0.8, False, train_box,
                              reduce_loss=False, end_points=ep, upstream=D(up))
    torch.cuda.synchronize()
    assert np.allclose(loss.cpu().numpy(), oloss.detach().numpy(), rtol=1e-4, atol=1e-6)
    z = lambda t: np.zeros_like(t.detach().numpy()) if t.grad is None else t.grad.numpy()
    ref_box = np.concatenate([z(oc), z(od), z(oo)[:, None]], axis=1)
    got_box = ep['surface_grad_box_reg'].cpu().numpy()
    scale_g = np.abs(ref_box).max() + 1e-9
    assert np.abs(got_box - ref_box).max() <= 2e-3 * scale_g, (got_box, ref_box)
    for k, on in zip((slice(0, 3), slice(3, 6), slice(6, 7)), train_box):
        if not on:
            assert np.all(got_box[:, k] == 0.0)
    assert np.allclose(ep['surface_grad_soft_mask'].cpu().numpy(), osm.grad.numpy(), rtol=1e-4, atol=1e-7)
    # reduce_loss and forward-only call
    total = W.get_surface_loss((D(center), D(dims), D(orient)), D(pc), D(soft), margin, scale, 0.8, False, train_box)
    assert abs(float(total) - float(oloss.mean())) <= 1e-4 * max(1.0, abs(float(oloss.mean())))


def test_surface_loss_is_zero_on_the_surface_and_grows_with_distance():
    """Points on the +x face of an axis-aligned unit box: distance 0; pushed out along x by t: distance t (the ray through
    the centre hits the same face)."""
    from transferable3d_b200 import weak_losses as W
    N = 256
    rng = np.random.RandomState(0)
    yz = rng.uniform(-0.05, 0.05, (N, 2)).astype(np.float32)      # near the axis: the (uncleaned) side-face distances stay larger
    box = (torch.zeros(1, 3).cuda(), torch.ones(1, 3).cuda(), torch.zeros(1).cuda())
    ones = torch.ones(1, N).cuda()
    for t in (0.0, 0.25):
        x = np.full((N, 1), 0.5, np.float32) * (1 + 2 * t)            # scaled along the ray: (0.5, y, z) * (1 + 2t)
        pts = np.concatenate([x, yz[:, :1] * (1 + 2 * t), yz[:, 1:] * (1 + 2 * t)], axis=1)[None]
        loss = W.get_surface_loss(box, torch.as_tensor(pts).cuda(), ones, 0.0, 1.0, 0.8, False, (True, False, True), reduce_loss=False)
        want = np.linalg.norm(pts[0], axis=1) * (1 - 1 / (1 + 2 * t))
        assert abs(float(loss[0]) - float(want.mean())) < 2e-4, (t, float(loss[0]), float(want.mean()))


def test_inactive_volume_loss_vs_oracle():
    """weak_losses.get_inactive_volume_loss_v1 (weak_losses.py:38-67): class-grouped mean violations, empty and untrained
    classes, all-satisfied margins."""
    from transferable3d_b200 import weak_losses as W
    from oracle import weak_losses as OW
    rng = np.random.RandomState(2)
    B = 300
    dims = rng.uniform(0.3, 2.0, (B, 3)).astype(np.float32)
    cls = rng.randint(0, 10, B)
    cls[cls == 7] = 6                                  # class 7 empty
    margins = np.array([4.0, 1.5, 2.0, 0.4, 0.3, 1.0, 0.8, 0.3, 1.0, 0.0], np.float32)      # class 9: never violated
    for train in ([True] * 10, [i % 2 == 0 for i in range(10)], [False] * 9 + [True]):
        want = OW.get_inactive_volume_loss_v1(torch.as_tensor(dims, dtype=torch.float64), torch.as_tensor(cls), train, 10,
                                              torch.as_tensor(margins, dtype=torch.float64))
        got = W.get_inactive_volume_loss_v1(torch.as_tensor(dims).cuda(), torch.as_tensor(cls).cuda(), train, 10, margins)
        assert abs(float(got) - float(want)) <= 1e-5 * max(1.0, abs(float(want))), (train, float(got), float(want))
