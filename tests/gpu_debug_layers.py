"""Debug helper (not a test): checks every TrainLayer.backward of the semisup-adv step against torch fp64 ops on the
same device inputs, to localise a faulty kernel / shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from test_gpu_semisup_train import _setup
from transferable3d_b200 import train_semisup_adv as tsa, train_layers as tl
B, N = int(sys.argv[1]), int(sys.argv[2])
v, feed, masks, FLAGS = _setup(B, N)
orig = tl.TrainLayer.backward

def rel(a, b):
    b = b.double(); a = a.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)), float((a - b).abs().mean() / b.abs().mean().clamp_min(1e-30))

def checked(self, dout, need_dx=True):
    d = dout.clone().double(); M = self.y.shape[0]
    W = self.W().double()
    if self.bn:
        out = self.out.double(); y = self.y.double(); mean = self.mean.double(); rstd = self.rstd.double()
        if self.act == 1: d = d * (out > 0)
        elif self.act == 2: d = d * torch.where(out > 0, 1.0, 0.2)
        elif self.act == 3: d = d * (1 - out * out)
        xh = (y - mean) * rstd
        s1 = d.sum(0); s2 = (d * xh).sum(0)
        dY = self.p('bn/gamma').double() * rstd * (d - s1 / M - xh * s2 / M)
    else:
        dY = d
    dW = self.x.double().t() @ dY
    dX = dY @ W.t()
    r = orig(self, dout, need_dx)
    torch.cuda.synchronize()
    msg = '%-45s M=%d K=%d N=%d' % (self.name, M, self.K, self.N)
    if self.bn:
        msg += ' beta %.1e/%.1e gamma %.1e/%.1e' % (rel(self.grads[self.name + '/bn/beta'], s1) + rel(self.grads[self.name + '/bn/gamma'], s2))
    msg += ' dW %.1e/%.1e' % rel(self.grads[self.name + '/weights'].view(self.K, self.N), dW)
    if r is not None:
        msg += ' dX %.1e/%.1e' % rel(r, dX)
    print(msg)
    return r
tl.TrainLayer.backward = checked
g = tsa.SemiAdvTrainGraph(v, FLAGS, B, N, 6, 'cuda:0')
ep = g.forward_backward(feed, masks)
torch.cuda.synchronize()
