"""Debug helper (not a test): per-variable gradient errors of the semisup-adv step vs the fp32 / fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from util import err_stats
from test_gpu_semisup_train import _setup
from transferable3d_b200 import train_semisup_adv as tsa
from oracle import train_semisup_adv as ot
B, N = int(sys.argv[1]), int(sys.argv[2])
v, feed, masks, FLAGS = _setup(B, N)
ol, og, ovs, oep = ot.loss_and_grads(v, FLAGS, feed, masks, 0, extra_grads=('stage1_center', 'feats_lv1'))
ol64, og64, _, oep64 = ot.loss_and_grads(v, FLAGS, feed, masks, 0, dtype=torch.float64, extra_grads=('stage1_center', 'feats_lv1'))
g = tsa.SemiAdvTrainGraph(v, FLAGS, B, N, 6, 'cuda:0')
ep = g.forward_backward(feed, masks)
torch.cuda.synchronize()
m = ep['mask'].cpu() > 0.5
m32 = oep['logits'][:, :, 0] < oep['logits'][:, :, 1]
m64 = oep64['logits'][:, :, 0] < oep64['logits'][:, :, 1]
print('mask mismatches gpu-vs-32', int((m != m32).sum()), 'gpu-vs-64', int((m != m64).sum()), '32-vs-64', int((m32 != m64).sum()))
print('loss', ep['loss_terms'].cpu().numpy(), float(ol), float(ol64))
for k in ('stage1_center', 'feats_lv1'):
    got = ep['d_' + k].cpu().numpy(); r64 = oep64['d_' + k].numpy(); r32 = oep['d_' + k].numpy()
    s = err_stats(got, r64); f = err_stats(r32, r64)
    print('d_%s gpu mean/scale %.2e max/scale %.2e | o32 %.2e %.2e' % (k, s['mean_abs'] / s['ref_scale'], s['max_abs'] / s['ref_scale'], f['mean_abs'] / f['ref_scale'], f['max_abs'] / f['ref_scale']))
    if k == 'stage1_center':
        print(got[:4]); print(r64[:4])
for k in og:
    if og[k] is None:
        continue
    got = g.grad[k].cpu().numpy().reshape(-1)
    s = err_stats(got, og64[k].numpy().reshape(-1)); f = err_stats(og[k].numpy().reshape(-1), og64[k].numpy().reshape(-1))
    print('%-55s gpu mean/scale %.2e max/scale %.2e | o32 mean/scale %.2e max/scale %.2e' % (k, s['mean_abs'] / max(s['ref_scale'], 1e-12),
          s['max_abs'] / max(s['ref_scale'], 1e-12), f['mean_abs'] / max(f['ref_scale'], 1e-12), f['max_abs'] / max(f['ref_scale'], 1e-12)))
print('---- per-channel view of box_est/conv-reg3')
k = 'class_agnostic/box_est/conv-reg3/bn/beta'
got = g.grad[k].cpu().numpy(); r64 = og64[k].numpy(); r32 = og[k].numpy()
err = np.abs(got - r64); idx = np.argsort(-err)[:8]
L = g.box[2]
frac = (L.out > 0).float().mean(dim=0).cpu().numpy()
print('worst channels', idx, 'err', err[idx], 'ref', r64[idx], 'o32err', np.abs(r32 - r64)[idx])
print('rstd', L.rstd.cpu().numpy()[idx], 'mean', L.mean.cpu().numpy()[idx], 'active frac', frac[idx])
print('median err', np.median(err), 'median rstd', np.median(L.rstd.cpu().numpy()))
kk = 'class_agnostic/box_est/conv-reg3/weights'
gw = g.grad[kk].cpu().numpy().reshape(128, 256); rw = og64[kk].numpy().reshape(128, 256)
ce = np.abs(gw - rw).max(axis=0); print('weights: worst cols', np.argsort(-ce)[:8], np.sort(-ce)[:8])
print('---- forward accuracy')
for k in ('stage1_center', 'feats_lv1', 'F_output', 'logits'):
    got = ep[k].cpu().numpy().astype(np.float64); r64 = oep64[k].detach().numpy(); r32 = oep[k].detach().numpy().astype(np.float64)
    print('%-16s gpu-vs-64 max %.2e mean %.2e | o32-vs-64 max %.2e mean %.2e | scale %.3f' % (k, np.abs(got - r64).max(), np.abs(got - r64).mean(),
          np.abs(r32 - r64).max(), np.abs(r32 - r64).mean(), np.abs(r64).mean()))
print('---- box_est backward recomputed with torch fp64 autograd on the GPU from our own inputs')
dev = 'cuda:0'
pc = torch.as_tensor(feed['pc']).to(dev)
s1c = ep['stage1_center'].double()
x0 = (pc[:, :, :3].double() - s1c[:, None, :]).reshape(B * N, 3).requires_grad_(True)
rowmask = ep['mask'].reshape(B * N).double()
Ws = []
x = x0
for l in g.box[:4]:
    W = l.W().double().clone().requires_grad_(True); b = l.p('biases').double(); ga = l.p('bn/gamma').double(); be = l.p('bn/beta').double()
    y = x @ W + b
    mean = y.mean(0); var = y.var(0, unbiased=False)
    x = torch.relu(ga * (y - mean) / torch.sqrt(var + 1e-3) + be)
    Ws.append(W)
pooled = (x * rowmask[:, None]).view(B, N, 512).max(1).values
print('pooled vs ours', float((pooled - ep['feats_lv1'].double()).abs().max()))
pooled.backward(ep['d_feats_lv1'].double())
for l, W in zip(g.box[:4], Ws):
    got = g.grad[l.name + '/weights'].view(l.K, l.N).double()
    print(l.name, 'dW ours-vs-torch64 max rel %.2e mean rel %.2e' % (float((got - W.grad).abs().max() / W.grad.abs().max()), float((got - W.grad).abs().mean() / W.grad.abs().mean())),
          '| oracle64-vs-torch64 mean rel %.2e' % float((torch.as_tensor(og64[l.name + '/weights'].numpy().reshape(l.K, l.N)).to(dev) - W.grad).abs().mean() / W.grad.abs().mean()))
gs = -x0.grad.view(B, N, 3).sum(1)
print('sum dX torch64', gs[:2].cpu().numpy())
