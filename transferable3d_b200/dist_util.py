"""Multi-GPU plumbing of the hot path: one process per GPU, torch.distributed for the rendezvous.

Inference shards frustums contiguously over the ranks with NO data-path collective (frustums are independent in eval
mode).  The training steps (train_boxpc / train_semisup_adv) are data parallel: every replica holds the full weights and
Adam state, and the only exchange per step is ONE sum all-reduce of the flat fp32 gradient arena, divided by the world
size inside the fused Adam kernel.  BN statistics, dropout masks and batch-coupled loss terms are per replica (the
reference is single-device; parity target = one replica at its local batch).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [begin, end) of `total` frustums owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError('rank %d outside world %d' % (rank, world))
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def world_size(pg=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(pg)
    return 1


def allreduce_flat(flat, pg=None):
    """In-place SUM all-reduce of one contiguous buffer (the gradient arena).  Returns the world size, by which the
    consumer scales (t3d_adam's grad_scale)."""
    w = world_size(pg)
    if w > 1:
        if not flat.is_contiguous():
            raise ValueError('the gradient arena must be one contiguous buffer')
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=pg)
    return w


def broadcast_params(flat, src=0, pg=None):
    """Replicas start from rank `src`'s parameters (what restoring one checkpoint on every replica does)."""
    if world_size(pg) > 1:
        dist.broadcast(flat, src=src, group=pg)


def average_moving_stats(tensors, pg=None):
    """Optional: average the BN moving statistics of the replicas (they drift apart because batch statistics are per
    replica).  One flat all-reduce."""
    w = world_size(pg)
    if w == 1 or not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=pg)
    flat /= w
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
