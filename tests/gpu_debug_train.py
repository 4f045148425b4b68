"""Not a test: layer-by-layer comparison of the BoxPC training forward against a CPU restatement."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from transferable3d_b200 import train_boxpc as tb, weights, synth, config
from oracle import train_boxpc as otb, tf_util as otu, boxpc_sunrgbd as obp
from util import err_stats

B, N = 8, 256
v = weights.make_weights_boxpc()
feed = synth.make_boxpc_batch(B, N, 6, seed=3)
rng = np.random.RandomState(3)
masks = {'dp1': (rng.rand(B, 512) < 0.7).astype(np.float32), 'dp2': (rng.rand(B, 256) < 0.7).astype(np.float32)}
FLAGS = config.cfg(BOXPC_WEIGHT_DELTA=4.)
g = tb.BoxPCTrainGraph(v, FLAGS, B, N, 6, 'cuda:0')
out = g.forward_backward(feed, masks)
torch.cuda.synchronize()
T = lambda a, dt=torch.float32: torch.as_tensor(np.asarray(a)).to(dt)
one_hot = T(feed['one_hot'])
x_box = (T(feed['x_center']), T(feed['x_orient_cls'], torch.int64), T(feed['x_orient_reg']), T(feed['x_dims_cls'], torch.int64), T(feed['x_dims_reg']))
box_reg = obp.convert_raw_y_box_to_reg_format(x_box, one_hot)
rep = otu.tf_get_box_pc_representation(box_reg, T(feed['pc'])).reshape(B * N, 12)
print('rep', err_stats(g.layers[0].x.cpu().numpy(), rep.numpy()))
x = rep
names = ['conv-reg1', 'conv-reg2', 'conv-reg3', 'conv-reg4', 'fc1', 'fc2', 'fc3']
for i, nm in enumerate(names):
    p = 'box_pc_mask_model/' + nm
    w = T(v[p + '/weights']).reshape(-1, v[p + '/weights'].shape[-1])
    if i == 4:
        x = x.reshape(B, N, 512).max(dim=1).values
        print('pooled', err_stats(g.layers[4].x.cpu().numpy(), x.numpy()))
    if i == 5:
        x = x * T(masks['dp1']) / 0.7
    if i == 6:
        x = x * T(masks['dp2']) / 0.7
    y = x @ w + T(v[p + '/biases'])
    print(nm, 'y  ', err_stats(g.layers[i].y.cpu().numpy(), y.numpy()))
    if i < 6:
        mu = y.mean(0); var = ((y - mu) ** 2).mean(0)
        y = torch.relu((y - mu) / torch.sqrt(var + 1e-3) * T(v[p + '/bn/gamma']) + T(v[p + '/bn/beta']))
        print(nm, 'out', err_stats(g.layers[i].out.cpu().numpy(), y.numpy()))
    x = y
print('out9', x[:2], out['output'][:2])
oloss, ograds, ovs, oep = otb.loss_and_grads(v, FLAGS, feed, masks)
print('oracle loss', float(oloss), 'gpu loss', float(out['loss']))
labels = (T(feed['y_box_iou']), (T(feed['y_center_delta']), T(feed['y_dims_delta']), T(feed['y_orient_delta'])))
pred = (x[:, -2:], (x[:, 0:3], x[:, 3:6], x[:, 6]))
l = obp.get_loss(pred, labels, {'logits_for_weigh': None}, c=FLAGS)
print('loss from restated out9', float(l))
print('cls', obp.get_boxpc_cls_loss(pred[0], labels[0], {}, False, FLAGS)[:4], out['boxpc_cls_losses'][:4])
print('delta', obp.get_boxpc_delta_loss(pred, labels, {'logits_for_weigh': None}, False, FLAGS)[:4], out['boxpc_delta_losses'][:4])
