"""bench.py -- frustums/sec of the Frustum-PointNet hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps K --warmup W
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference algorithm (oracle, PyTorch-CPU) on the host cores

Workload (config.workload): BASELINE.json cfg3 = full Frustum PointNet v1 pipeline inference
(seg -> mask / centroid / resample 512 -> T-Net -> box-est NH=12 NS=10), 8192 frustums of
2048 pts x 6 ch + 10-class one-hot.  Frustums are independent, so the GPUs shard them with no data-path
collective; `--scaling weak` (default) gives every GPU the full batch of 8192, `--scaling strong` splits
one global batch of 8192 contiguously.  `--workload cfg2` times the instance-seg chain alone at batch 1024 per GPU.
A step = one pass of the pipeline over the rank's shard, processed in chunks of --chunk frustums.
`value`: inputs already resident in HBM.  `e2e`: the same pass through the public API
(frustum_pointnets_v1.get_model) from pinned HOST buffers, H2D of every chunk's inputs and D2H of
its results (mask logits + box outputs) inside the timed region, copies overlapped with compute.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOTAL_FRUSTUMS = {'cfg3': 8192, 'cfg2': 1024, 'cfg1': 32, 'cfg4': 256, 'cfg5': 256}
N_POINTS, N_CH = 2048, 6
UNIQUE = 8192           # unique synthetic frustums generated on the host, tiled to the workload size
# algorithmic FLOPs (2*MAC) per point, SURVEY 8(d) / DESIGN.md
FLOP_SEG1_PT = 295680
FLOP_SEG2_PT = 426496
FLOP_SEG_GLOBAL_FR = 2 * (1024 + 10) * 512
FLOP_TNET_PT, FLOP_BOX_PT = 99072, 361216
FLOP_FC_FR = 2 * (98688 + 2560 + 410368 + 5120)
# DRAM bytes per frustum (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture at 8192 frustums per
# launch of the final round-2 build, profiles/r02_ncu_cfg3_kernels_final.csv); algorithmic: stage 2 reads point_feat 262 144 B +
# gbias 2 048 B and writes logits 16 384 B, stage 1 reads the points 49 152 B and writes point_feat 262 144 B + gfeat 4 096 B
NCU_DRAM_BYTES_FR = {'seg2': (2.169e9 + 0.134e9) / 8192, 'seg1': (0.437e9 + 2.124e9) / 8192}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d['bf16_tflops_sustained'], bf16_burst=d['bf16_tflops'], hbm=d['hbm_gbs'], src='measured')
    return dict(bf16=1400.0, bf16_burst=1590.0, hbm=6650.0, src='fallback')


class ClockSampler(object):
    """SM clock / power / clock-event reasons sampled DURING the timed region: NVML polled every 5 ms from a
    thread of this process (a timed region of a few steps lasts tens of ms, too short for `nvidia-smi -lms 200`);
    falls back to the profiling recipe's nvidia-smi loop if NVML is unavailable."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.samples, self._stop = None, [], threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(',')) else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)),
                                     int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h)),
                                     n.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=1)
            n = self.nvml
            names = (('hw_slowdown', n.nvmlClocksEventReasonHwSlowdown), ('hw_thermal_slowdown', n.nvmlClocksEventReasonHwThermalSlowdown),
                     ('sw_thermal_slowdown', n.nvmlClocksEventReasonSwThermalSlowdown), ('sw_power_cap', n.nvmlClocksEventReasonSwPowerCap))
            reasons = sorted({nm for _, r, _ in self.samples for nm, bit in names if r & bit})
            sm = [x[0] for x in self.samples]
            pw = [x[2] for x in self.samples]
            return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': self.sm_max, 'reasons': reasons,
                    'samples': len(sm), 'power_w_max': max(pw) if pw else None, 'source': 'nvml, 5 ms poll during the timed region'}
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'source': 'nvidia-smi -lms 200'}


def make_host_data(workload, n_local, seed, wire=False):
    """Synthetic frustums of the workload.  Colours are quantised to 8 bits, rgb = float32(k) / 255 -- what im2double of an
    8-bit image gives in the prepared SUN-RGBD frustums -- so that the e2e wire format (xyz fp32 + rgb uint8) and the
    fp32 (B,N,6) placeholder hold bit-identical values.  wire=True also returns (xyz [n,N,3] fp32, rgb [n,N,3] uint8)."""
    from transferable3d_b200 import synth
    u = min(UNIQUE, n_local)
    b = synth.make_batch(u, N_POINTS, N_CH, seed=seed)
    k = np.clip(np.rint(b['pc'][:, :, 3:6] * 255.0), 0, 255).astype(np.uint8)
    pc_u = b['pc'].copy()
    pc_u[:, :, 3:6] = k.astype(np.float32) / np.float32(255)
    reps = (n_local + u - 1) // u
    pc = np.ascontiguousarray(np.tile(pc_u, (reps, 1, 1))[:n_local])
    oh = np.ascontiguousarray(np.tile(b['one_hot'], (reps, 1))[:n_local])
    if not wire:
        return pc, oh
    return pc, oh, np.ascontiguousarray(pc[:, :, 0:3]), np.ascontiguousarray(np.tile(k, (reps, 1, 1))[:n_local])


def standard_variables(workload):
    from transferable3d_b200 import weights
    if workload == 'cfg3':
        return weights.standard_model_A()
    v, info = weights.standard_model_F()
    return v, info


# ------------------------------------------------------------------------------------------ reference arm (CPU)

def run_reference(args):
    """The reference algorithm (oracle restatement of the TF1 graph, literal tiled-global conv6, every
    activation materialised) on the host cores; PyTorch-CPU with all threads.  Not TensorFlow."""
    import torch
    from oracle.tf_layers import VarStore
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    workload = args.workload
    torch.set_num_threads(os.cpu_count() or 1)
    if workload in ('cfg4', 'cfg5'):
        from oracle import train_boxpc as otb, train_semisup_adv as ota
        sample = 8
        v2, feed2, masks2, FLAGS2 = train_setup(workload, sample, N_POINTS, 77)
        feed2 = {k: a for k, a in feed2.items() if k != '_2d'}
        fn = otb.loss_and_grads if workload == 'cfg4' else ota.loss_and_grads

        def step():
            fn(v2, FLAGS2, feed2, masks2)
    else:
        variables, _ = standard_variables(workload)
        sample = 32 if workload == 'cfg1' else args.ref_sample
        pc, oh = make_host_data(workload, sample, 1234 + (3 if workload == 'cfg3' else 2))
        vs = VarStore(variables)
        vs.literal = True
        pc_t, oh_t = torch.as_tensor(pc), torch.as_tensor(oh)

        def step():
            with torch.no_grad():
                if workload == 'cfg3':
                    oracle_cfg3(vs, pc_t, oh_t)
                elif workload == 'cfg1':
                    from oracle import test_semisup as ots
                    from transferable3d_b200 import config
                    ots.run_graph(vs, config.cfg(), pc_t, oh_t)
                else:
                    from oracle import semisup_models as osm
                    with vs.variable_scope('class_agnostic'):
                        osm.v1_inst_seg(pc_t, None, None, {}, False, vs, scope='inst_seg')
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    line = {'impl': 'reference', 'metric': 'frustums_per_sec', 'value': value, 'unit': 'frustums/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
            'scaling': args.scaling if workload == 'cfg3' else 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(workload, args, sample_note='reference arm: %d frustums per step' % sample),
            'cpu_baseline': {'value': value, 'unit': 'frustums/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                             'sample': '%d frustums per step, oracle restatement of the TF1 graph on PyTorch-CPU' % sample},
            'e2e': {'value': value, 'unit': 'frustums/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def oracle_cfg3(vs, pc_t, oh_t, seed=5):
    """SURVEY 3.2 pipeline on the oracle (model-A variable names): oracle/frustum_pointnets_v1.py."""
    from oracle import frustum_pointnets_v1 as ofpn
    return ofpn.get_model(vs, pc_t, oh_t, seed=seed)


def workload_config(workload, args, sample_note=None):
    c = {'workload': ('cfg3: Frustum PointNet v1 pipeline inference (seg -> mask/centroid/resample 512 -> T-Net -> '
                      'box-est NH=12 NS=10), 8192 frustums x 2048 pts x 6 ch + one-hot %s, frustums sharded over the GPUs, '
                      'no collective' % ('per GPU' if args.scaling == 'weak' else 'in total'))
         if workload == 'cfg3' else {'cfg1': 'cfg1: semisup_v1_sunrgbd model F + 1 BoxPC refine, eval, batch 32 x 2048 pts x 6 ch',
                                     'cfg2': 'cfg2: instance-seg per-point MLP chain alone, 1024 frustums x 2048 pts x 6 ch per GPU',
                                     'cfg4': 'cfg4: train_boxpc BoxPC-Fit forward + backward, batch 256 x 2048 pts per GPU',
                                     'cfg5': 'cfg5: train_semisup_adv model F training step, batch 256 x 2048 pts per GPU'}[workload],
         'global_frustums': TOTAL_FRUSTUMS[workload] * (1 if (workload == 'cfg3' and args.scaling == 'strong') else args.gpus),
         'frustums_per_gpu': TOTAL_FRUSTUMS[workload] // (args.gpus if (workload == 'cfg3' and args.scaling == 'strong') else 1),
         'num_point': N_POINTS, 'num_channel': N_CH, 'chunk_frustums': args.resident_chunk, 'e2e_chunk_frustums': args.chunk, 'parallelism': 'shard%d' % args.gpus,
         'precision': {'bf16': 'bf16 operands / fp32 accumulate (tcgen05), fp32 heads',
                       'f16x2': 'fp16 hi + lo operand pairs, 3 tcgen05 products per layer, fp32 accumulate (fp32-accurate), fp32 heads'}[getattr(args, 'precision', 'bf16')],
         'weights': 'synthetic Xavier (seed 42), seg logits calibrated (margin std 2.0, 40% masked-in)',
         'frustums': '%d unique synthetic frustums per rank, colours quantised to 8 bit (k / 255)' % UNIQUE,
         'resample_rng': 'philox', 'l2': 'per-step inputs (>= 400 MB per GPU at N=1) exceed the 126 MB L2; no explicit flush'}
    if sample_note:
        c['sample'] = sample_note
    return c


# ------------------------------------------------------------------------------------------ cfg1 (the reference's own batch-32 case)

def run_cfg1(args):
    """BASELINE cfg1: semisup_v1_sunrgbd model F + one BoxPC refine, eval mode, batch 32, through test_semisup.get_model /
    sess.run with the fetch list of test_semisup.inference (test_semisup.py:210-226).  A step = one batch of 32 frustums.
    value: inputs resident on the device, CUDA-graph replay; e2e: numpy feed in, fetched tensors back on the host."""
    import torch
    from transferable3d_b200 import weights, config, runtime as rt, test_semisup as ts
    assert args.gpus == 1 and int(os.environ.get('WORLD_SIZE', '1')) == 1, 'cfg1 is a single-GPU latency case'
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    B = TOTAL_FRUSTUMS['cfg1']
    variables, _ = weights.standard_model_F()
    FLAGS = config.cfg()
    rt.set_precision('bf16')
    pc_h, oh_h = make_host_data('cfg1', B, 1235)
    fetch = ['logits', 'F2_center', 'F2_heading_scores', 'F2_heading_residuals', 'F2_size_scores', 'F2_size_residuals', 'boxpc_fit_prob']
    res = {}
    for name, graph in (('graph', True), ('eager', False)):
        sess, ops = ts.get_model(B, N_POINTS, N_CH, FLAGS=FLAGS, variables=variables, device=dev, cuda_graph=graph)
        pc_d, oh_d = torch.as_tensor(pc_h).to(dev), torch.as_tensor(oh_h).to(dev)
        feed_d = {ops['pc_pl']: pc_d, ops['one_hot_vec_pl']: oh_d, ops['is_training_pl']: False}
        feed_h = {ops['pc_pl']: pc_h, ops['one_hot_vec_pl']: oh_h, ops['is_training_pl']: False}

        def timed(fn):
            for _ in range(max(args.warmup, 3)):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(args.steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / args.steps
        ms = timed(lambda: sess.run(fetch, feed_d))
        ms_e2e = timed(lambda: [t.cpu() for t in sess.run(fetch, feed_h)])
        res[name] = (ms, ms_e2e)
    ms, ms_e2e = res['graph']
    d2h = sum(int(np.prod(s_)) * 4 for s_ in ((B, N_POINTS, 2), (B, 3), (B, 12), (B, 12), (B, 10), (B, 10, 3), (B,)))
    flops_fr = (FLOP_SEG1_PT + FLOP_SEG2_PT + FLOP_TNET_PT + FLOP_BOX_PT + 363520) * N_POINTS   # dense upper bound: masked stacks run compacted
    line = {'metric': 'frustums_per_sec', 'value': B / ms * 1e3, 'unit': 'frustums/s', 'n_gpus': 1, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': 'cfg1: semisup_v1_sunrgbd model F + 1 BoxPC refine, eval, batch 32 x 2048 pts x 6 ch, '
                                   'test_semisup.get_model / sess.run, CUDA-graph replay', 'batch': B, 'num_point': N_POINTS,
                       'l2': 'latency case: 1.5 MB of inputs per step, L2-resident by nature'},
            'e2e': {'value': B / ms_e2e * 1e3, 'unit': 'frustums/s', 'h2d_bytes_per_step': pc_h.nbytes + oh_h.nbytes,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e},
            'eager_launch_path': {'ms_per_step': res['eager'][0], 'e2e_ms_per_step': res['eager'][1]},
            'gpu_launches': None, 'roofline': None, 'dense_tflops_upper_bound': B / ms * 1e3 * flops_fr / 1e12}
    if not args.no_cpu_baseline:
        from oracle.tf_layers import VarStore
        from oracle import test_semisup as ots
        torch.set_num_threads(os.cpu_count() or 1)
        vs = VarStore(variables)
        vs.literal = True
        pc_t, oh_t = torch.as_tensor(pc_h), torch.as_tensor(oh_h)
        with torch.no_grad():
            ots.run_graph(vs, FLAGS, pc_t, oh_t)
            t0, reps = time.perf_counter(), 0
            while reps < 3 or time.perf_counter() - t0 < 10.0:
                ots.run_graph(vs, FLAGS, pc_t, oh_t)
                reps += 1
        dt = (time.perf_counter() - t0) / reps
        line['cpu_baseline'] = {'value': B / dt, 'unit': 'frustums/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                                'sample': '%d batches of 32, oracle restatement of the TF1 graph (literal conv6) on PyTorch-CPU' % reps}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ cfg4 / cfg5 (training steps)

CFG5_FLAGS = dict(SEMI_TRAIN_BOX_TRAIN_CLASS_AG_TNET=True, SEMI_TRAIN_BOX_TRAIN_CLASS_AG_BOX=True, SEMI_BOXPC_MIN_FIT_LOSS_AFT_REFINE=True,
                  WEAK_WEIGHT_INTRACLASSVAR=2., WEAK_WEIGHT_REPROJECTION=0.01, WEAK_REPROJECTION_ONLY_ON_2D_CLS=True,
                  SEMI_MULTIPLIER_FOR_WEAK_LOSS=0.05, SEMI_WEIGHT_BOXPC_FIT_LOSS=1.)


def train_setup(workload, B, N, seed):
    from transferable3d_b200 import weights, synth, config
    rng = np.random.RandomState(seed)
    if workload == 'cfg4':
        v = weights.make_weights_boxpc()
        feed = synth.make_boxpc_batch(B, N, N_CH, seed=seed)
        masks = {'dp1': (rng.rand(B, 512) < 0.7).astype(np.float32), 'dp2': (rng.rand(B, 256) < 0.7).astype(np.float32)}
        return v, feed, masks, config.cfg(BOXPC_WEIGHT_DELTA=4.)
    v = weights.make_weights_model_F()
    # pure batches alternating 3D-labelled / 2D-only (ALTERNATE_BATCH, train_semisup_adv.py:539-565); 2D batches
    # carry all-zero 3D labels (roi_semi_dataset.py:452-454).  `feed` is the 3D batch, feed['_2d'] the 2D one.
    feed = synth.make_batch(B, N, N_CH, seed=seed, is_data_2D=0)
    feed2 = synth.make_batch(B, N, N_CH, seed=seed + 1, is_data_2D=1)
    for k in ('labels', 'centers', 'y_orient_cls', 'y_orient_reg', 'y_dims_cls', 'y_dims_reg'):
        if k in feed2:
            feed2[k] = np.zeros_like(feed2[k])
    feed['_2d'] = feed2
    masks = {'class_agnostic/inst_seg/dp1': (rng.rand(B, N, 128) < 0.5).astype(np.float32),
             'class_dependent/box_refine/dp0': (rng.rand(B, 512) < 0.5).astype(np.float32),
             'class_dependent/box_refine/dp1': (rng.rand(B, 256) < 0.5).astype(np.float32)}
    return v, feed, masks, config.cfg(**CFG5_FLAGS)


def run_train(args):
    """BASELINE cfg4 (train_boxpc: BoxPC-Fit forward + backward + Adam) / cfg5 (train_semisup_adv: model F training step with
    the frozen BoxPC branch and the strong + reprojection + intra-class-variance + fit losses), batch 256 per GPU, data
    parallel with ONE NCCL all-reduce of the flat gradient arena per step (weak scaling).  A step = one training step;
    value = frustums/s with the batch resident on the device; e2e = numpy batch in (H2D inside) and the loss read back."""
    import torch
    import torch.distributed as dist
    from transferable3d_b200 import train_boxpc as tb, train_semisup_adv as tsa
    rank, world, local = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    assert world == args.gpus
    workload, B, N = args.workload, TOTAL_FRUSTUMS[args.workload], N_POINTS
    from transferable3d_b200 import runtime as rt
    rt.set_f32_engine(args.f32_engine)
    v, feed, masks, FLAGS = train_setup(workload, B, N, 1234 + (4 if workload == 'cfg4' else 5) + 100 * rank)
    g = tb.BoxPCTrainGraph(v, FLAGS, B, N, N_CH, dev) if workload == 'cfg4' else tsa.SemiAdvTrainGraph(v, FLAGS, B, N, N_CH, dev)
    if world > 1:      # identical replicas to start from
        from transferable3d_b200.dist_util import broadcast_params
        broadcast_params(g.flat_param if workload == 'cfg4' else g.arena.flat_param)
    D = lambda a: torch.as_tensor(np.asarray(a)).to(dev)
    feeds_h = [feed]
    if '_2d' in feed:                              # cfg5: alternate the pure 3D batch and the pure 2D batch
        feeds_h = [{k: a for k, a in feed.items() if k != '_2d'}, feed['_2d']]
    feeds_d = [{k: D(a) for k, a in f.items()} for f in feeds_h]
    masks_d = {k: D(a) for k, a in masks.items()}
    loss_key = 'loss' if workload == 'cfg4' else 'semi_loss'
    losses, nstep = [], {'n': 0}

    def step_resident():
        out = g.step(feeds_d[nstep['n'] % len(feeds_d)], masks_d)
        nstep['n'] += 1
        losses.append(out[loss_key])

    def step_e2e():
        out = g.step(feeds_h[nstep['n'] % len(feeds_h)], masks_d)   # numpy batch -> device inside the step; dropout masks are device state
        nstep['n'] += 1
        losses.append(float(out[loss_key].reshape(-1)[0]))

    def timed(fn, sample_clocks=False):
        for _ in range(max(args.warmup, 3)):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks
    from transferable3d_b200 import _lib as _l, train_layers as _tl, losses as _ls
    import transferable3d_b200.tf_util as _tu
    launches = {'n': 0, 'on': False}
    _orig = _l.call

    def _counting(name, *a):
        if launches['on']:
            launches['n'] += 1
        return _orig(name, *a)
    for mod in (_l, _tl, _ls, _tu, tb, tsa, rt):
        if hasattr(mod, 'call'):
            mod.call = _counting
    ms, clocks = timed(step_resident, True)
    launches['on'] = True
    step_resident()                                   # C-ABI calls of one step (each launches >= 1 kernel of this library)
    torch.cuda.synchronize()
    launches['on'] = False
    n_launch = launches['n']
    first, last = float(losses[0].reshape(-1)[0]), float(losses[-1].reshape(-1)[0])
    ms_e2e, _ = timed(step_e2e)
    h2d = sum(np.asarray(a).nbytes for a in feeds_h[0].values())
    # algorithmic FLOPs of one step (2 x MAC; SURVEY 8(d): dgrad = wgrad = forward MACs per trained layer, forward only for
    # frozen layers, no dgrad into data).  Per-point / per-frustum MAC counts of the layer tables.
    P = N
    if workload == 'cfg4':      # BoxPC net: conv 12-128-128-256-512 + fc 512-512-256-9, all trained, no dgrad for conv1
        macs_fr = 3 * (181760 * P + 395520) - 12 * 128 * P
    else:                       # seg forward (frozen, conv6 folded) + T-Net and box convs trained (box FC head forward only)
        seg = 361088 * P + 1024 * 512                                  # conv1-5, conv6' (K = 64), conv7-10 per point + global half
        tnet = 3 * (49536 * P + 98688) - 3 * 128 * P                    # no dgrad into the points
        box = 3 * (180608 * P) + 410368                                 # conv1 dgrad feeds stage1_center
        refine = 3 * 415488
        boxpc = 2 * (181760 * P + 395520)                               # frozen branch: forward + dgrad to the box
        macs_fr = seg + tnet + box + refine + boxpc
    step_flops = 2.0 * macs_fr * B
    pk = peaks()
    roof = {'kernel': 'whole training step (tcgen05 bf16 x 3 GEMMs with fused BN statistics / lazy BN + pooling / loss / Adam kernels)', 'bound': 'tensor',
            'achieved': step_flops / (ms * 1e-3) / 1e12, 'peak': pk['bf16_burst'], 'unit': 'TFLOP/s',
            'frac': step_flops / (ms * 1e-3) / 1e12 / pk['bf16_burst'], 'traffic': None,
            'algorithmic_flops_per_step': step_flops,
            'note': 'GEMMs issue %s bf16 tensor-core products per MAC (engine %s); batch-norm statistics come out of '
                    'the GEMM epilogues and the BN map is applied by the consumers (lazy batch norm); the step is bound by the '
                    'loaders / epilogues of the fp32-operand GEMM kernels and the fp32 activation traffic, not by the tensor pipe '
                    '(DESIGN.md section 5)' % ({'tc': 6, 'tc2': 3, 'bf16': 1, 'simt': 0}[args.f32_engine], args.f32_engine),
            'peak_source': '%s, burst bf16' % pk['src']}
    nparam = int((g.flat_param if workload == 'cfg4' else g.arena.flat_param).numel())
    line = {'metric': 'frustums_per_sec', 'value': B * world / ms * 1e3, 'unit': 'frustums/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': ('cfg4: train_boxpc BoxPC-Fit (rep A) forward + backward + Adam' if workload == 'cfg4' else
                                    'cfg5: train_semisup_adv model F step (reprojection + intra-class-variance + fit losses, frozen BoxPC)') +
                       ', batch %d x %d pts per GPU, data parallel, one NCCL all-reduce of the flat fp32 gradient arena per step' % (B, N),
                       'global_batch': B * world, 'allreduce_bytes': 4 * nparam, 'parallelism': 'dp%d' % world,
                       'precision': {'tc': 'fp32 in HBM; GEMMs on tcgen05 with every operand split into 3 bf16 pieces, 6 partial products, '
                                           'fp32 accumulate (fp32-accurate); BN / loss / Adam kernels fp32 on CUDA cores',
                                     'tc2': 'fp32 in HBM; GEMMs on tcgen05 with every operand split into 2 bf16 pieces (rounded to nearest, '
                                            'residual <= 2^-18), 3 partial products, fp32 accumulate: ~1e-5 of sum|a||b| (between TF32 and fp32); '
                                            'BN / loss / Adam kernels fp32 on CUDA cores',
                                     'simt': 'fp32 CUDA-core kernels throughout',
                                     'bf16': 'fp32 in HBM; GEMM operands rounded to bf16, one tcgen05 pass, fp32 accumulate; '
                                             'BN / loss / Adam kernels fp32'}[args.f32_engine],
                       'f32_engine': args.f32_engine,
                       'l2': 'activations of one step (> 2 GB) exceed the 126 MB L2'},
            'clocks': clocks, 'e2e': {'value': B * world / ms_e2e * 1e3, 'unit': 'frustums/s', 'h2d_bytes_per_step': h2d,
                                      'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e},
            'loss_first_step': first, 'loss_last_step': last, 'gpu_launches': n_launch, 'roofline': roof}
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            from oracle import train_boxpc as otb, train_semisup_adv as ota
            torch.set_num_threads(os.cpu_count() or 1)
            Bs = 8
            v2, feed2, masks2, FLAGS2 = train_setup(workload, Bs, N, 77)
            feed2 = {k: a for k, a in feed2.items() if k != '_2d'}
            fn = otb.loss_and_grads if workload == 'cfg4' else ota.loss_and_grads
            fn(v2, FLAGS2, feed2, masks2)
            t0, reps = time.perf_counter(), 0
            while reps < 2 or time.perf_counter() - t0 < 10.0:
                fn(v2, FLAGS2, feed2, masks2)
                reps += 1
            dt = (time.perf_counter() - t0) / reps
            line['cpu_baseline'] = {'value': Bs / dt, 'unit': 'frustums/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                                    'sample': '%d steps of batch %d (forward + autograd backward, no optimizer), oracle restatement on PyTorch-CPU' % (reps, Bs)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ B200 arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='t3d', choices=['t3d', 'reference'])
    ap.add_argument('--workload', default='cfg3', choices=['cfg3', 'cfg2', 'cfg1', 'cfg4', 'cfg5'])
    ap.add_argument('--chunk', type=int, default=8192, help='frustums per chunk of the e2e (H2D/compute/D2H) pipeline')
    ap.add_argument('--resident-chunk', type=int, default=8192, help='frustums per pass when inputs are resident in HBM')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: every GPU processes the full cfg3 batch of 8192 frustums (no data-path collective); '
                         'strong: one global batch of 8192 split over the GPUs')
    ap.add_argument('--f32-engine', default='tc', choices=['tc', 'tc2', 'simt', 'bf16'],
                    help='GEMM engine of the training workloads (cfg4 / cfg5): tc = tcgen05 bf16 x 3 split (fp32-accurate), '
                         'tc2 = bf16 x 2 split (three products, ~1e-5), '
                         'simt = CUDA-core SGEMM, bf16 = one tcgen05 pass on bf16-rounded operands')
    ap.add_argument('--ref-sample', type=int, default=32, help='frustums per step of the reference (CPU) arm')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'f16x2'],
                    help='mode of the headline line: bf16 = one tensor-core product per MAC; f16x2 = fp16 hi / lo split, three products, '
                         'fp32-accurate (the mode that meets the mask-exactness target); the other one is reported under exact_mode')
    ap.add_argument('--no-exact-mode', action='store_true', help='skip the f16x2 leg of the line')
    ap.add_argument('--no-parity', action='store_true', help='skip the oracle spot-check of the timed batch')
    ap.add_argument('--cpu-sample', type=int, default=8)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--breakdown', action='store_true', help='per-entry-point CUDA-event times of the resident step (stderr)')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
        return
    if args.workload == 'cfg1':
        run_cfg1(args)
        return
    if args.workload in ('cfg4', 'cfg5'):
        run_train(args)
        return

    import torch
    import torch.distributed as dist
    from transferable3d_b200 import runtime as rt, _lib, model_util as mu, semisup_models as sm
    from transferable3d_b200 import frustum_pointnets_v1 as fpn

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the B200 path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    assert world == args.gpus, 'launch with torchrun --nproc-per-node %d' % args.gpus
    workload = args.workload
    total = TOTAL_FRUSTUMS[workload]
    n_local = total // world if (workload == 'cfg3' and args.scaling == 'strong') else total
    chunk = min(args.chunk, n_local)
    rchunk = min(args.resident_chunk, n_local)
    assert n_local % chunk == 0 and n_local % rchunk == 0

    variables, winfo = standard_variables(workload)
    store = rt.VariableStore(variables, dev)
    rt.set_default_store(store)
    rt.set_precision(args.precision)
    mu.set_resample_rng('philox', seed=5)
    numa = bind_numa_local(local)
    pc_h, oh_h, xyz_h, rgb_h = make_host_data(workload, n_local, 1234 + (3 if workload == 'cfg3' else 2), wire=True)
    xyz_pin, rgb_pin, oh_pin = torch.from_numpy(xyz_h).pin_memory(), torch.from_numpy(rgb_h).pin_memory(), torch.from_numpy(oh_h).pin_memory()
    pc_dev, oh_dev = torch.from_numpy(pc_h).to(dev), oh_pin.to(dev)

    # launch counting + per-kernel timing of the two dominant kernels (CUDA events on the launching stream)
    launches = {'n': 0}
    KERNELS_PER_CALL = {'t3d_chain_max_bf16': 1, 't3d_seg_stage2_bf16': 1, 't3d_chain_max_x2': 1, 't3d_seg_stage2_x2': 1,
                        't3d_linear_f32': 1, 't3d_mask_centroid': 1, 't3d_resample': 1, 't3d_build_tiles': 1, 't3d_parse_box': 1,
                        't3d_prepare_xyz': 1, 't3d_assemble_points': 1}
    dom = {'seg2': [], 'seg1': [], 'on': False}
    orig_call = _lib.call
    bd = {'on': False, 'ev': []}

    def counting_call(name, *a):
        launches['n'] += KERNELS_PER_CALL.get(name, 0)
        if name == 't3d_linear_f32' and a[9] >= 4096 and a[11] >= 64 and 32 <= a[10] <= 2048 and rt.get_f32_engine() != 'simt':
            launches['n'] += 1          # the weight pre-split pass of the tensor-core GEMM (xg_presplit_kernel)
        if bd['on']:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig_call(name, *a)
            e1.record()
            bd['ev'].append((name + (':%d' % a[0] if name.startswith('t3d_chain_max') else ''), e0, e1))
            return
        is_seg2 = name in ('t3d_seg_stage2_bf16', 't3d_seg_stage2_x2')
        is_seg1 = name in ('t3d_chain_max_bf16', 't3d_chain_max_x2') and a[0] == 0
        if dom['on'] and (is_seg2 or is_seg1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig_call(name, *a)
            e1.record()
            dom['seg2' if is_seg2 else 'seg1'].append((e0, e1))
        else:
            orig_call(name, *a)
    import transferable3d_b200.tf_util as tu
    import transferable3d_b200.boxpc_sunrgbd as bp
    import transferable3d_b200.test_semisup as ts
    for mod in (rt, sm, mu, _lib, tu, bp, ts):
        if hasattr(mod, 'call'):
            mod.call = counting_call

    def pipeline(pc, oh):
        if workload == 'cfg3':
            return fpn.get_model(pc, oh, False)
        return {'mask_logits': sm.v1_inst_seg(pc, None, None, {}, False, scope='class_agnostic/inst_seg')}

    def step_resident():
        with torch.no_grad():
            for c0 in range(0, n_local, rchunk):
                pipeline(pc_dev[c0:c0 + rchunk], oh_dev[c0:c0 + rchunk])

    # e2e through the public API a user calls: host wire format (xyz fp32 + rgb uint8 + one-hot, pinned) -> H2D ->
    # model_util.assemble_point_cloud(lazy=True) -> frustum_pointnets_v1.inference (pipeline + the reference's test-time
    # post-processing on the device) -> D2H of the prediction (pred_seg bytes + 10 numbers per frustum).  Double-buffered:
    # inputs land in two preallocated device buffer sets (copy stream), the pipeline runs on the compute stream, results are
    # packed into two preallocated staging sets (D2D) and leave for pinned host memory on a third stream -- no allocation
    # and no record_stream inside the timed region.  cfg2 (seg chain alone) returns the raw mask logits.
    OUT_KEYS = ('pred_seg', 'center', 'heading_cls', 'heading_res', 'size_cls', 'size_res', 'scores') if workload == 'cfg3' else ('mask_logits',)
    copy_s, back_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    in_bufs = [(torch.empty((chunk, N_POINTS, 3), device=dev), torch.empty((chunk, N_POINTS, 3), dtype=torch.uint8, device=dev),
                torch.empty((chunk, 10), device=dev)) for _ in range(2)]
    in_bytes = sum(t.numel() * t.element_size() for t in in_bufs[0][:3])
    stage_out, host_out = [{}, {}], [{}, {}]
    bytes_io = {'h2d': 0, 'd2h': 0}
    # a stream of batches: chunk g (numbered over all steps) lives in buffer set g % 2; the H2D of chunk g + 1 -- the next step's
    # first chunk when g closes a step -- is issued before chunk g is computed, D2H copies drain under the next chunk's compute.
    # Every step issues exactly n_local / chunk input copies inside the timed region (the first timed step consumes one copy
    # issued by the last warm-up step and the last timed step issues one for the step after it).
    carry = {'g': 0, 'ready': {}, 'consumed': [None, None], 'drained': [None, None]}

    def e2e_pipeline(s):
        xyz, rgb, oh = in_bufs[s]
        pc = mu.assemble_point_cloud(xyz, rgb, lazy=True)      # bf16: the inst_seg chain reads the wire format directly
        if workload == 'cfg3':
            launches['n'] += 1                      # t3d_inference_scores (bound outside _lib.call)
            return fpn.inference(pc, oh)
        return pipeline(pc, oh)

    def step_e2e():
        comp = torch.cuda.current_stream()
        bytes_io['h2d'] = bytes_io['d2h'] = 0
        nchunks = n_local // chunk
        ready = carry['ready']
        consumed = carry['consumed']
        drained = carry['drained']      # the D2H of the staging set has finished

        def issue_copy(g):
            s, i = g % 2, g % nchunks
            with torch.cuda.stream(copy_s):
                if consumed[s] is not None:
                    copy_s.wait_event(consumed[s])
                sl = slice(i * chunk, (i + 1) * chunk)
                in_bufs[s][0].copy_(xyz_pin[sl], non_blocking=True)
                in_bufs[s][1].copy_(rgb_pin[sl], non_blocking=True)
                in_bufs[s][2].copy_(oh_pin[sl], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_s)
                ready[g] = ev
            bytes_io['h2d'] += in_bytes
        with torch.no_grad():
            for _ in range(nchunks):
                g = carry['g']
                s = g % 2
                if g not in ready:
                    issue_copy(g)            # very first chunk of the stream only
                comp.wait_event(ready.pop(g))
                ep = e2e_pipeline(s)
                ev = torch.cuda.Event()
                ev.record(comp)
                consumed[s] = ev
                issue_copy(g + 1)            # into the other buffer set: waits (on the copy stream) for the chunk that used it last
                carry['g'] = g + 1
                if drained[s] is not None:
                    comp.wait_event(drained[s])
                for k in OUT_KEYS:
                    t = ep[k]
                    if k not in stage_out[s]:
                        stage_out[s][k] = torch.empty_like(t)
                        host_out[s][k] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
                    stage_out[s][k].copy_(t)
                ev2 = torch.cuda.Event()
                ev2.record(comp)
                with torch.cuda.stream(back_s):
                    back_s.wait_event(ev2)
                    for k in OUT_KEYS:
                        host_out[s][k].copy_(stage_out[s][k], non_blocking=True)
                        bytes_io['d2h'] += stage_out[s][k].numel() * stage_out[s][k].element_size()
                    e3 = torch.cuda.Event()
                    e3.record(back_s)
                    drained[s] = e3

    def e2e_finish():
        """before the closing timestamp: every result of every timed step is on the host"""
        comp = torch.cuda.current_stream()
        for e in carry['drained']:
            if e is not None:
                comp.wait_event(e)

    def timed(fn, steps, warmup, sample_clocks=False, finish=None):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        launches['n'] = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches['n'], clocks

    pk = peaks()
    n_calls = args.steps * (n_local // rchunk)

    def measure_resident(precision):
        """resident step in `precision`: ms per step, launches, clocks and the rooflines of the two dominant kernels"""
        rt.set_precision(precision)
        dom['seg2'], dom['seg1'], dom['on'] = [], [], True
        ms, n_launch, clocks = timed(step_resident, args.steps, args.warmup, sample_clocks=True)
        dom['on'] = False
        region_s = ms * args.steps * 1e-3
        # the driver-measured bf16 peak a kernel is held to: the burst figure when the timed region is shorter than the 4 s
        # sustained-clock measurement window, as here (steps x ~12 ms); the sustained figure rides along
        peak, peak_name = (pk['bf16_burst'], 'burst') if region_s < 1.0 else (pk['bf16'], 'sustained')
        products = 3 if precision == 'f16x2' else 1
        roofs = {}
        for key, flop_pt, kname, traffic, traffic_alg in (
                ('seg2', FLOP_SEG2_PT, 'seg_stage2_x2_kernel' if products == 3 else 'seg_stage2_pair_kernel',
                 NCU_DRAM_BYTES_FR['seg2'], 262144 + 2048 + 16384),
                ('seg1', FLOP_SEG1_PT, 'chain_max_x2_kernel<SEG1>' if products == 3 else 'chain_max_kernel<SEG1>',
                 NCU_DRAM_BYTES_FR['seg1'], 49152 + 262144 + 4096)):
            t_ms = [a.elapsed_time(b) for a, b in dom[key][-n_calls:]]
            if not t_ms:
                continue
            avg = float(np.mean(t_ms))
            flops = flop_pt * rchunk * N_POINTS
            ach = flops / (avg * 1e-3) / 1e12
            r = {'kernel': kname + (' (conv6..conv10, tcgen05)' if key == 'seg2' else ' (conv1..conv5 + max-pool fused, tcgen05)'),
                 'bound': 'tensor', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak,
                 'frac_of_sustained_peak': ach / pk['bf16'], 'frac_of_burst_peak': ach / pk['bf16_burst'],
                 'avg_launch_ms': avg, 'algorithmic_flops_per_launch': flops,
                 'peak_source': '%s, %s bf16 (timed region %.2f s)' % (pk['src'], peak_name, region_s)}
            if products == 1:
                r.update({'traffic': traffic * rchunk, 'traffic_algorithmic': traffic_alg * rchunk,
                          'traffic_source': 'dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel at '
                                            '8192 frustums per launch (profiles/), scaled per frustum; not re-measured in this run'})
            else:
                # three tensor-core products per algorithmic MAC (hi.lo + lo.hi + hi.hi): what the tensor pipe executes
                r.update({'traffic': None, 'executed_tflops': 3 * ach, 'executed_frac_of_peak': 3 * ach / peak,
                          'note': 'f16x2: fp32-accurate result from 3 fp16 tensor-core products per MAC; achieved / frac count the '
                                  'algorithmic FLOPs once, executed_* count what the tensor pipe runs'})
            roofs[key] = r
        return ms, n_launch, clocks, roofs

    ms, n_launch, clocks, roofs = measure_resident(args.precision)
    ms_e2e, n_launch_e2e, _ = timed(step_e2e, args.steps, args.warmup, finish=e2e_finish)
    exact = None
    if args.precision != 'f16x2' and not args.no_exact_mode:
        ms_x, _, clocks_x, roofs_x = measure_resident('f16x2')
        exact = {'precision': 'f16x2', 'value': n_local * world / ms_x * 1e3, 'unit': 'frustums/s', 'ms_per_step': ms_x,
                 'roofline': roofs_x.get('seg2'), 'roofline_fused_maxpool': roofs_x.get('seg1'), 'clocks': clocks_x}
        rt.set_precision(args.precision)

    if args.breakdown and rank == 0:
        bd['on'] = True
        step_resident()
        torch.cuda.synchronize()
        bd['on'] = False
        agg = {}
        for name, a, b in bd['ev']:
            agg.setdefault(name, [0, 0.0])
            agg[name][0] += 1
            agg[name][1] += a.elapsed_time(b)
        tot = sum(v[1] for v in agg.values())
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            sys.stderr.write('  %-28s n=%3d  %8.3f ms  %.3f\n' % (name, n, t, t / tot))
    total_units = n_local * world
    value = total_units / ms * 1e3
    e2e_value = total_units / ms_e2e * 1e3
    if workload == 'cfg3':
        flops_fr = (FLOP_SEG1_PT + FLOP_SEG2_PT) * N_POINTS + FLOP_SEG_GLOBAL_FR + (FLOP_TNET_PT + FLOP_BOX_PT) * 512 + FLOP_FC_FR
    else:
        flops_fr = (FLOP_SEG1_PT + FLOP_SEG2_PT) * N_POINTS + 2 * 1024 * 512
    line = {'metric': 'frustums_per_sec', 'value': value, 'unit': 'frustums/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': args.scaling if workload == 'cfg3' else 'weak', 'vs_baseline': None,
            'dtype': {'bf16': 'bf16', 'f16x2': 'f16 x 2 (fp32-accurate)'}[args.precision], 'data': 'synthetic',
            'config': workload_config(workload, args), 'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'frustums/s', 'h2d_bytes_per_step': bytes_io['h2d'],
                    'd2h_bytes_per_step': bytes_io['d2h'], 'ms_per_step': ms_e2e, 'gpu_launches': n_launch_e2e,
                    'host_gbps_all_ranks': (bytes_io['h2d'] + bytes_io['d2h']) * world / (ms_e2e * 1e-3) / 1e9,
                    'wire_format': 'in: xyz fp32 + rgb uint8 + one-hot fp32 (pinned); out: pred_seg uint8 + centre / heading / size / score'
                                   if workload == 'cfg3' else 'in: xyz fp32 + rgb uint8; out: raw mask logits fp32', 'numa': numa},
            'gpu_launches': n_launch, 'roofline': roofs.get('seg2'), 'roofline_fused_maxpool': roofs.get('seg1'),
            'pipeline_tflops': value * flops_fr / 1e12, 'pipeline_frac_of_bf16_peak': value * flops_fr / 1e12 / (pk['bf16_burst'] * world),
            'exact_mode': exact}

    if rank == 0:
        if not args.no_parity:
            line['parity'] = {args.precision: parity_block(workload, variables, pc_h, oh_h, pc_dev, oh_dev, rchunk, args.precision)}
            if exact is not None:
                line['parity']['f16x2'] = parity_block(workload, variables, pc_h, oh_h, pc_dev, oh_dev, rchunk, 'f16x2')
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline(workload, variables, args.cpu_sample)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bind_numa_local(local_rank):
    """Pin this rank to the CPUs NVML reports as local to its GPU BEFORE the pinned host buffers are allocated (first touch
    puts them on that NUMA node).  Returns what was done, for the bench line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        phys = int(vis.split(',')[local_rank]) if vis and all(x.strip().isdigit() for x in vis.split(',')) else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        node = None
        try:
            node = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
        return {'gpu': phys, 'cpus_local_to_gpu': '%d-%d (%d)' % (min(allowed), max(allowed), len(allowed)) if allowed else None,
                'numa_node': node}
    except Exception as e:      # NVML missing: leave the affinity alone
        return {'error': str(e)[:80]}


def parity_block(workload, variables, pc_h, oh_h, pc_dev, oh_dev, rchunk, precision, n_sample=16):
    """Checker leg (not timed): 16 frustums sampled from the timed batch against the oracle.  Seg logits and masks first;
    then, with the oracle continued from the GPU's own logits (identical masks, identical resampled indices), every box
    output.  The sample is re-run as its own batch (the Philox key of the resample is (seed, index in the batch)) and
    tied to the timed batch by comparing its logits with the big batch's."""
    import torch
    from oracle.tf_layers import VarStore
    from transferable3d_b200 import runtime as rt, frustum_pointnets_v1 as fpn, semisup_models as sm
    sel = np.sort(np.random.RandomState(7).permutation(rchunk)[:n_sample])
    torch.set_num_threads(os.cpu_count() or 1)
    vs = VarStore(variables)
    vs.literal = True
    pc_t, oh_t = torch.as_tensor(pc_h[sel]), torch.as_tensor(oh_h[sel])
    rt.set_precision(precision)
    with torch.no_grad():
        if workload == 'cfg3':
            from oracle import frustum_pointnets_v1 as ofpn
            big = fpn.get_model(pc_dev[:rchunk], oh_dev[:rchunk], False)['mask_logits'][torch.as_tensor(sel, device=pc_dev.device)].cpu().numpy()
            ep = fpn.get_model(pc_dev[:rchunk][torch.as_tensor(sel, device=pc_dev.device)].contiguous(),
                               oh_dev[:rchunk][torch.as_tensor(sel, device=pc_dev.device)].contiguous(), False)
            g = ep['mask_logits'].cpu().numpy()
            oep = ofpn.get_model(vs, pc_t, oh_t, seed=5)
            ol = oep['mask_logits'].numpy()
            cont = ofpn.get_model(vs, pc_t, oh_t, seed=5, logits=torch.as_tensor(g))
        else:
            from oracle import semisup_models as osm
            big = sm.v1_inst_seg(pc_dev[:rchunk], None, None, {}, False, scope='class_agnostic/inst_seg')[torch.as_tensor(sel, device=pc_dev.device)].cpu().numpy()
            g, ep, cont = big, None, None
            with vs.variable_scope('class_agnostic'):
                ol = osm.v1_inst_seg(pc_t, None, None, {}, False, vs, scope='inst_seg').numpy()
    scale = float(np.abs(ol).mean())
    agree = (g[..., 0] < g[..., 1]) == (ol[..., 0] < ol[..., 1])
    out = {'precision': precision, 'frustums_checked': int(n_sample), 'oracle': 'PyTorch-CPU fp32 restatement of the TF1 graph (literal conv6)',
           'seg_logit_err_max_of_scale': float(np.abs(g - ol).max() / scale), 'seg_logit_err_mean_of_scale': float(np.abs(g - ol).mean() / scale),
           'mask_point_agreement': float(agree.mean()), 'mask_frustums_bit_exact': '%d / %d' % (int(agree.all(axis=1).sum()), n_sample),
           'mask_exactness_1024_frustums': 'profiles/r02_mask_exactness.txt',
           'timed_batch_vs_checked_batch_logit_diff_of_scale': float(np.abs(big - g).max() / scale)}
    if cont is not None:
        out['resampled_indices_bit_exact'] = bool(np.array_equal(ep['object_pc_indices'].cpu().numpy(), cont['object_pc_indices']))
        worst, inside = 0.0, 1.0
        for k in ('stage1_center', 'center', 'heading_scores', 'heading_residuals', 'size_scores', 'size_residuals'):
            a, b = ep[k].float().cpu().numpy().astype(np.float64), cont[k].numpy().astype(np.float64)
            worst = max(worst, float(np.abs(a - b).max()))
            inside = min(inside, float((np.abs(a - b) <= 1e-3 + 1e-2 * np.abs(b)).mean()))
        out['box_outputs_from_gpu_logits'] = {'max_abs_err': worst, 'min_fraction_within_rel1e-2_abs1e-3': inside}
    return out


def cpu_baseline(workload, variables, sample):
    """Oracle (literal TF1-graph restatement, PyTorch-CPU fp32, all host threads) on a bounded sample."""
    import torch
    from oracle.tf_layers import VarStore
    torch.set_num_threads(os.cpu_count() or 1)
    pc, oh = make_host_data(workload, sample, 99)
    vs = VarStore(variables)
    vs.literal = True
    pc_t, oh_t = torch.as_tensor(pc), torch.as_tensor(oh)

    def step():
        with torch.no_grad():
            if workload == 'cfg3':
                oracle_cfg3(vs, pc_t, oh_t)
            else:
                from oracle import semisup_models as osm
                with vs.variable_scope('class_agnostic'):
                    osm.v1_inst_seg(pc_t, None, None, {}, False, vs, scope='inst_seg')
    step()
    t0 = time.perf_counter()
    reps = 0
    while reps < 3 or time.perf_counter() - t0 < 10.0:
        step()
        reps += 1
        if time.perf_counter() - t0 > 30.0:
            break
    dt = (time.perf_counter() - t0) / reps
    return {'value': sample / dt, 'unit': 'frustums/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': '%d frustums x %d reps, oracle restatement of the TF1 graph (literal conv6) on PyTorch-CPU' % (sample, reps)}


if __name__ == '__main__':
    main()
