"""-m gpu: model-A semi-supervised training step (train_semisup.py: get_semi_model_backbone in training mode,
get_semi_loss_backbone, Adam over ALL variables -- the segmentation network trains) against the oracle (PyTorch autograd
restatement + TF's Adam rule) on the same seeded batch and dropout mask."""
import numpy as np
import pytest
import torch

from util import err_stats, assert_grad_close

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from transferable3d_b200 import train_semisup as ts, weights, synth, config

DEV = 'cuda:0'
CFG_A = dict(SEMI_MODEL='A', WEAK_WEIGHT_REPROJECTION=0.01, SEMI_MULTIPLIER_FOR_WEAK_LOSS=0.05)


def _setup(B, N, seed=11, **over):
    v = weights.make_weights_model_A()
    feed = synth.make_batch(B, N, 6, seed=seed, is_data_2D=(np.arange(B) % 2))
    rng = np.random.RandomState(seed)
    masks = {'inst_seg/dp1': (rng.rand(B, N, 128) < 0.5).astype(np.float32)}
    kw = dict(CFG_A)
    kw.update(over)
    return v, feed, masks, config.cfg(**kw)


@pytest.mark.parametrize('B,N,over', [(8, 256, {}), (16, 512, {}), (8, 256, dict(WEAK_WEIGHT_SURFACE=0.)),
                                      (8, 256, dict(WEAK_TRAIN_BOX_W_SURFACE=(True, False, True), WEAK_REPROJECTION_USE_SOFTMAX_PROJ=True))])
def test_semisup_model_a_step_vs_oracle(B, N, over, built_lib):
    from oracle import train_semisup as ot
    v, feed, masks, FLAGS = _setup(B, N, **over)
    oloss, ograds, ovs, oep = ot.loss_and_grads(v, FLAGS, feed, masks, global_step=0)
    _, ograds64, _, oep64 = ot.loss_and_grads(v, FLAGS, feed, masks, global_step=0, dtype=torch.float64)
    g = ts.SemiTrainGraph(v, FLAGS, B, N, 6, DEV)
    ep = g.forward_backward(feed, masks)
    torch.cuda.synchronize()
    omask = (oep['logits'][:, :, 0] < oep['logits'][:, :, 1])
    omask64 = (oep64['logits'][:, :, 0] < oep64['logits'][:, :, 1])
    stable = (omask == omask64)
    assert torch.equal(ep['mask'].cpu()[stable] > 0.5, omask[stable])
    s = err_stats(ep['logits'].cpu().numpy(), oep['logits'].detach().numpy())
    assert s['max_abs'] <= 1e-3 * max(s['ref_scale'], 1.0), s
    for k in ('stage1_center', 'center', 'heading_scores', 'size_residuals', 'soft_mask'):
        s = err_stats(ep[k].cpu().numpy(), oep[k].detach().numpy())
        assert s['max_abs'] <= 2e-3 * max(s['ref_scale'], 1.0), (k, s)
    assert abs(float(ep['semi_loss']) - float(oloss)) <= 2e-4 * max(1.0, abs(float(oloss))), (float(ep['semi_loss']), float(oloss))
    # every variable has a gradient in the oracle (no var_list): compare all of them
    wscale = {}
    assert set(ograds) == set(g.grad)
    for name, og in ograds.items():
        assert og is not None, name
        ref, ref64 = og.numpy().reshape(-1), ograds64[name].numpy().reshape(-1)
        layer = name.rsplit('/', 1)[0] if not name.endswith(('gamma', 'beta')) else name.rsplit('/', 2)[0]
        if name.endswith('weights'):
            wscale[layer] = float(np.abs(ref64).mean())
        assert_grad_close(name, g.grad[name].cpu().numpy(), ref, ref64, scale_floor=1e-2 * wscale.get(layer, 0.0))
    for k, mv in g.moving.items():
        ref = ovs.vars[k].numpy()
        small = k.endswith('variance') and '/fc' in k
        s = err_stats(mv.cpu().numpy(), ref)
        assert s['max_abs'] <= (2e-2 if small else 5e-4) * max(s['ref_scale'], 1e-3), (k, s)
    # one TF-Adam update
    from oracle.train_boxpc import adam_step_tf, get_learning_rate
    g.apply_gradients()
    lr = get_learning_rate(0, B)
    for name in ('inst_seg/conv6/weights', 'tnet/fc3-stage1/weights', 'box_est/fc3/biases', 'inst_seg/conv1/bn/gamma'):
        p0 = torch.as_tensor(v[name]).reshape(-1)
        ref, _, _ = adam_step_tf(p0, g.grad[name].cpu().reshape(-1), torch.zeros_like(p0), torch.zeros_like(p0), lr, 1)
        assert torch.allclose(g.param[name].cpu(), ref, atol=1e-6, rtol=1e-5), name
    assert g.global_step == 1


def test_semisup_model_a_training_reduces_loss(built_lib):
    v, feed, masks, FLAGS = _setup(8, 256)
    g = ts.SemiTrainGraph(v, FLAGS, 8, 256, 6, DEV)
    losses = [float(g.step(feed, masks)['semi_loss']) for _ in range(12)]
    assert np.isfinite(losses).all() and losses[-1] < 0.8 * losses[0], losses
