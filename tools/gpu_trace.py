"""Not a test: per-role clock64 timeline of CTA 0 of the tcgen05 kernels (t3d_set_trace_buffer),
plus H2D/D2H bandwidth of this box.  Run under gpurun."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transferable3d_b200 import weights, synth, runtime as rt, _lib, semisup_models as sm  # noqa: E402

SLOTS = 8192


def dump(buf, title, tiles=(6, 7)):
    b = buf.cpu().numpy().reshape(4, SLOTS)
    print('=== %s' % title)
    for role, name in enumerate(('producer', 'mma', 'epilogue', 'front')):
        ev = b[role]
        ev = ev[ev != 0]
        if len(ev) == 0:
            continue
        t = (ev >> 8).astype(np.int64)
        tag = (ev & 0xff).astype(np.int64)
        # tile boundaries = tag 0x10 (producer has none: print stats only)
        if role == 0:
            d = np.diff(t)
            print('  %-9s %d events, mean gap %.0f cyc, p95 %.0f' % (name, len(ev), d.mean(), np.percentile(d, 95)))
            continue
        starts = np.where(tag == 0x10)[0]
        print('  %-9s %d events, %d tiles, cycles/tile median %.0f' % (
            name, len(ev), len(starts), np.median(np.diff(t[starts])) if len(starts) > 1 else -1))
        for ti in tiles:
            if ti + 1 >= len(starts):
                continue
            s, e = starts[ti], starts[ti + 1]
            rel = t[s:e + 1] - t[s]
            print('     tile %d: ' % ti + ' '.join('%02x@%d' % (tg, r) for tg, r in zip(tag[s:e + 1], rel)))


def main():
    dev = 'cuda:0'
    variables, _ = weights.standard_model_F()
    store = rt.VariableStore(variables, dev)
    rt.set_default_store(store)
    B = 148 * 2
    b = synth.make_batch(64, 2048, 6, seed=3)
    pc = torch.as_tensor(np.tile(b['pc'], (B // 64 + 1, 1, 1))[:B]).to(dev)
    with torch.no_grad():
        sm.v1_inst_seg(pc, None, None, {}, False, scope='class_agnostic/inst_seg')
        torch.cuda.synchronize()
        buf = torch.zeros(4 * SLOTS, dtype=torch.int64, device=dev)
        _lib.call('t3d_set_trace_buffer', _lib.ptr(buf))
        full = 'class_agnostic/inst_seg'
        arena1 = store.chain_arena(full, rt.CHAIN_SEG1, ['conv1', 'conv2', 'conv3', 'conv4', 'conv5'])
        pf = torch.empty((B * 2048, 64), dtype=torch.bfloat16, device=dev)
        g = rt.chain_max(rt.CHAIN_SEG1, pc, arena1, emit=pf)
        torch.cuda.synchronize()
        dump(buf, 'chain_max<SEG1> (256-pt tiles; tags: mma 2x=act ready l,3x=hidden issued,40=final ready,5x=acc empty mt,6x=issued mt;'
                  ' epi 2x=hidden full (l*2+sub),3x=hidden done,5x=final full,6x=final done; front 10=start,11=done)')
        buf.zero_()
        w6, b6 = store.folded(full + '/conv6')
        gbias, _ = rt.linear(g, w6[64:].contiguous(), b6)
        rt.seg_stage2(pf, gbias, store.seg2_arena(full), B, 2048)
        torch.cuda.synchronize()
        dump(buf, 'seg_stage2 (128-pt tiles; mma 2x/28+=job6 start/end, 3x/38+=job7, 40/41 conv8, 50/51 conv9; epi 2x/28+ e6, 30/31 e7, 40/41 e8, 50/51 e9)')
        buf.zero_()
        # box chain on 512 gathered-like points
        pc3 = pc[:, :512, :3].contiguous()
        arena_b = store.chain_arena('class_agnostic/box_est', rt.CHAIN_BOX, ['conv-reg1', 'conv-reg2', 'conv-reg3', 'conv-reg4'])
        rt.chain_max(rt.CHAIN_BOX, pc3, arena_b)
        torch.cuda.synchronize()
        dump(buf, 'chain_max<BOX> (128-pt tiles)')
        _lib.call('t3d_set_trace_buffer', None)
    # host <-> device bandwidth
    x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    y = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, fn in (('H2D', lambda: y.copy_(x, non_blocking=True)), ('D2H', lambda: x.copy_(y, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print('%s pinned 256 MiB: %.1f GB/s' % (name, 4 * 256 * 1.048576e6 / (e0.elapsed_time(e1) * 1e-3) / 1e9))


if __name__ == '__main__':
    main()
