"""Run under torchrun on >= 2 GPUs (not collected by pytest):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_train.py
Checks the data-parallel training steps (cfg4 BoxPC, cfg5 semi-supervised): every rank runs forward/backward on its own
micro-batch, ONE NCCL all-reduce of the flat gradient arena, fused Adam with 1/world scaling.  After the step the
parameters are identical on all ranks and equal to a single-process Adam update with the mean of the per-rank gradients."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from transferable3d_b200 import train_boxpc as tb, train_semisup_adv as tsa, weights, synth, config   # noqa: E402
from oracle import train_boxpc as otb                                                                  # noqa: E402  (checker)


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    B, N = 8, 256
    # ---- cfg4
    v = weights.make_weights_boxpc()
    feed = synth.make_boxpc_batch(B, N, 6, seed=100 + rank)
    rng = np.random.RandomState(rank)
    masks = {'dp1': (rng.rand(B, 512) < 0.7).astype(np.float32), 'dp2': (rng.rand(B, 256) < 0.7).astype(np.float32)}
    g = tb.BoxPCTrainGraph(v, config.cfg(BOXPC_WEIGHT_DELTA=4.), B, N, 6, dev)
    g.forward_backward(feed, masks)
    check_step(g.flat_param, g.flat_grad, lambda: g.apply_gradients(), otb.get_learning_rate(0, B), world, 'cfg4')
    # ---- cfg5
    from test_gpu_semisup_train import CFG5
    v = weights.make_weights_model_F()
    feed = synth.make_batch(B, N, 6, seed=200 + rank, is_data_2D=(np.arange(B) % 2))
    masks = {'class_agnostic/inst_seg/dp1': (rng.rand(B, N, 128) < 0.5).astype(np.float32),
             'class_dependent/box_refine/dp0': (rng.rand(B, 512) < 0.5).astype(np.float32),
             'class_dependent/box_refine/dp1': (rng.rand(B, 256) < 0.5).astype(np.float32)}
    g = tsa.SemiAdvTrainGraph(v, config.cfg(**CFG5), B, N, 6, dev)
    g.forward_backward(feed, masks)
    check_step(g.arena.flat_param, g.arena.flat_grad, lambda: g.apply_gradients(), otb.get_learning_rate(0, B), world, 'cfg5')
    dist.barrier()
    if rank == 0:
        print('multi_gpu_train ok: world %d, flat all-reduce + Adam identical on all ranks' % world)
    dist.destroy_process_group()


def check_step(flat_param, flat_grad, apply, lr, world, tag):
    local_grad = flat_grad.clone()
    p0 = flat_param.clone()
    gathered = [torch.empty_like(local_grad) for _ in range(world)]
    dist.all_gather(gathered, local_grad)
    mean_grad = torch.stack(gathered).sum(0) / world
    assert not torch.equal(gathered[0], gathered[1]), tag + ': ranks must see different micro-batches'
    apply()
    torch.cuda.synchronize()
    ref, _, _ = otb.adam_step_tf(p0.cpu(), mean_grad.cpu(), torch.zeros_like(p0.cpu()), torch.zeros_like(p0.cpu()), lr, 1)
    assert torch.allclose(flat_param.cpu(), ref, atol=1e-6, rtol=1e-5), tag + ': Adam on the averaged gradient'
    allp = [torch.empty_like(flat_param) for _ in range(world)]
    dist.all_gather(allp, flat_param)
    for q in allp[1:]:
        assert torch.equal(q, allp[0]), tag + ': replicas diverged'


if __name__ == '__main__':
    main()
