"""Not a test: prints per-stage error statistics of the B200 path against the oracle (used while
developing kernels; run under gpurun, output goes to stdout / gpurun_out)."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from transferable3d_b200 import weights, synth, config, runtime as rt  # noqa: E402
from transferable3d_b200 import semisup_models as sm, test_semisup as ts, model_util as mu  # noqa: E402
from transferable3d_b200 import frustum_pointnets_v1 as fpn  # noqa: E402
from oracle.tf_layers import VarStore  # noqa: E402
from oracle import semisup_models as osm, test_semisup as ots, model_util as omu  # noqa: E402
from util import err_stats  # noqa: E402


def show(name, got, ref):
    got = got.detach().float().cpu().numpy() if torch.is_tensor(got) else np.asarray(got)
    ref = ref.detach().float().cpu().numpy() if torch.is_tensor(ref) else np.asarray(ref)
    s = err_stats(got, ref)
    ok = np.abs(got - ref) <= 1e-3 + 1e-2 * np.abs(ref)
    print('  %-28s max_abs %.3e mean_abs %.3e scale %.3e  frac(rel1e-2/abs1e-3) %.5f  nan %d' % (
        name, s['max_abs'], s['mean_abs'], s['ref_scale'], ok.mean(), int(np.isnan(got).sum())), flush=True)


def section(title):
    print('\n=== %s' % title, flush=True)


def main():
    B = int(os.environ.get('DBG_B', '4'))
    dev = 'cuda:0'
    variables, info = weights.standard_model_F()
    print('weights', info)
    b = synth.make_batch(B, 2048, 6, seed=1234)
    FLAGS = config.cfg()
    vs = VarStore(variables)
    pc_c, oh_c = torch.as_tensor(b['pc']), torch.as_tensor(b['one_hot'])
    with torch.no_grad():
        ologits, oep = ots.run_graph(vs, FLAGS, pc_c, oh_c)
    store = rt.VariableStore(variables, dev)
    rt.set_default_store(store)
    pc, oh = pc_c.to(dev), oh_c.to(dev)

    def stage_checks(mode):
        section('per-stage, precision=%s (each stage fed the ORACLE inputs)' % mode)
        with rt.precision(mode), torch.no_grad():
            try:
                ep = {}
                with rt.variable_scope('class_agnostic'):
                    logits = sm.v1_inst_seg(pc, None, None, ep, False, scope='inst_seg')
                torch.cuda.synchronize()
                show('seg logits', logits, ologits)
                d = (logits[..., 1] - logits[..., 0]).cpu()
                od = ologits[..., 1] - ologits[..., 0]
                agree = ((d > 0) == (od > 0)).float()
                print('  mask per-point agreement %.5f  frustum-exact %.3f  oracle mask frac %.3f' % (
                    agree.mean(), (agree.min(dim=1).values).mean(), (od > 0).float().mean()))
            except Exception:
                traceback.print_exc()
            try:
                # downstream stages on the oracle's logits
                olog = ologits.to(dev).contiguous()
                mask, mean, xyz, xyz1 = sm.subtract_points_mean(pc, olog)
                omask = (ologits[..., 0:1] < ologits[..., 1:2]).float()
                print('  mask bit-exact on identical logits:', bool(torch.equal(mask.cpu(), omask)))
                ep = {}
                with rt.variable_scope('class_agnostic'):
                    s1 = sm.v1_tnet(xyz1, mask, mean, None, ep, False, scope='tnet')
                    show('stage1_center', s1, oep['stage1_center'])
                    os1 = oep['stage1_center'].to(dev)
                    sub = sm.subtract_1st_stage_center(xyz, os1)
                    sm.v1_box_est(sub, os1, mask, None, ep, False, scope='box_est')
                    show('feats_lv1', ep['feats_lv1'], oep['feats_lv1'])
                    show('box_params', ep['box_params'], oep['box_params'])
                    show('center', ep['center'], oep['center'])
            except Exception:
                traceback.print_exc()
            try:
                from transferable3d_b200 import boxpc_sunrgbd as bp
                obox = tuple(t.to(dev).contiguous() for t in oep['F_pred_box_reg'])
                with rt.variable_scope('D_boxpc_branch'):
                    pred, bep = bp.get_model((obox, pc), False, oh, use_one_hot_vec=False, c=FLAGS)
                show('boxpc feats_lv1', bep['boxpc_feats_dict']['box_pc_mask_model_feats_lv1'],
                     oep['boxpc_feats_dict']['box_pc_mask_model_feats_lv1'])
                show('boxpc delta_center', bep['boxpc_delta_center'], oep['boxpc_delta_center'])
                show('boxpc fit prob', bep['logits_for_weigh'], oep['boxpc_fit_prob'])
            except Exception:
                traceback.print_exc()

    def e2e(mode):
        section('end-to-end model F + 1 BoxPC refine, precision=%s' % mode)
        with rt.precision(mode), torch.no_grad():
            try:
                logits, ep = ts.build_graph(FLAGS, pc, oh)
                torch.cuda.synchronize()
                show('logits', logits, ologits)
                for k in ('stage1_center', 'F_center', 'F_heading_scores', 'F_heading_residuals', 'F_size_scores',
                          'F_size_residuals', 'F2_center', 'F2_heading_residuals', 'F2_size_residuals', 'boxpc_fit_prob'):
                    show(k, ep[k], oep[k])
            except Exception:
                traceback.print_exc()

    stage_checks('fp32')
    e2e('fp32')
    stage_checks('bf16')
    e2e('bf16')

    section('resampling (philox) bit-exactness on identical logits')
    try:
        olog = ologits.to(dev).contiguous()
        mu.set_resample_rng('philox', seed=99)
        ep = {}
        obj, mean, ep = mu.point_cloud_masking(pc, olog, ep)
        oepx = {}
        oobj, omean, oepx = omu.point_cloud_masking(pc_c, ologits, oepx, rng_mode='philox', seed=99)
        print('  indices bit-exact:', bool(np.array_equal(ep['object_pc_indices'].cpu().numpy(), oepx['object_pc_indices'])))
        show('object_pc', obj, oobj)
    except Exception:
        traceback.print_exc()

    section('timing (B=%d)' % int(os.environ.get('DBG_TB', '256')))
    try:
        TB = int(os.environ.get('DBG_TB', '256'))
        bb = synth.make_batch(TB, 2048, 6, seed=5)
        pcb, ohb = torch.as_tensor(bb['pc']).to(dev), torch.as_tensor(bb['one_hot']).to(dev)
        with rt.precision('bf16'), torch.no_grad():
            for name, fn in (('seg only', lambda: sm.v1_inst_seg(pcb, None, None, {}, False, scope='class_agnostic/inst_seg')),
                             ('model F + refine', lambda: ts.build_graph(FLAGS, pcb, ohb))):
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                print('  %-20s %.3f ms/iter  %.0f frustums/s' % (name, ms, TB / ms * 1e3), flush=True)
    except Exception:
        traceback.print_exc()


if __name__ == '__main__':
    main()
