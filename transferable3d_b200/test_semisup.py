"""Mirror of the inference side of sunrgbd/sunrgbd_detection/test_semisup.py: get_model (:61-179,
model F + BoxPC refine loop -> F2_* end points) and inference (:187-260).

The reference builds a static TF graph for a fixed batch size and returns (sess, ops); here
get_model returns (sess, ops) where `sess.run(fetches, feed_dict)` evaluates the same graph on
the B200 for whatever batch is fed (ops hold placeholder names).
"""
import numpy as np
import torch

from . import runtime as rt
from . import tf_util, semisup_v1_sunrgbd as MODEL, boxpc_sunrgbd
from ._lib import ptr, stream, call
from .constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER, NUM_CLASS


def build_graph(FLAGS, pc, one_hot_vec, box2D=None, img_dim=None, oracle_mask=None, is_training=False):
    """Body of test_semisup.get_model (test_semisup.py:74-149) on device tensors -> (logits, end_points)."""
    norm_box2D = None
    if box2D is not None and img_dim is not None:
        norm_box2D = tf_util.tf_normalize_2D_bboxes(box2D, img_dim)
    pred, end_points = MODEL.get_semi_model(pc, None, None, one_hot_vec, is_training, norm_box2D=norm_box2D,
                                            use_one_hot=FLAGS.use_one_hot, oracle_mask=oracle_mask, c=FLAGS)
    logits = pred[0]
    prefix = 'F_'
    n_refine = int(FLAGS.refine)
    B = pc.shape[0]
    dev = pc.device
    curr_box = tuple(t.clone() for t in end_points[prefix + 'pred_box_reg'])
    totals = (torch.zeros((B, 3), device=dev), torch.zeros((B, 3), device=dev), torch.zeros((B,), device=dev))
    boxpc_fit_prob = None
    for i in range(n_refine):
        if FLAGS.mask_pc_for_boxpc:
            mask = torch.argmax(logits, dim=2).to(torch.float32).unsqueeze(2)
            fake_pc = (pc * mask).contiguous()
        else:
            fake_pc = pc
        box_in = tuple(t.clone() for t in curr_box)
        with rt.variable_scope('D_boxpc_branch'):
            _, ep = boxpc_sunrgbd.get_model((box_in, fake_pc), False, one_hot_vec=one_hot_vec, use_one_hot_vec=False,
                                            c=FLAGS,
                                            # the test graph forces SEMI_WEIGH_BOXPC_DELTA_DURING_TEST = False whatever the
                                            # training configuration says (reference test_semisup.py:86)
                                            _refine=dict(curr_box=curr_box, totals=totals, weigh_during_test=False))
        boxpc_fit_prob = ep['logits_for_weigh']
        for k in ('boxpc_delta_center', 'boxpc_delta_size', 'boxpc_delta_angle', 'boxpc_feats_dict', 'pred_boxpc_fit',
                  'boxpc_fit_logits'):
            end_points[k] = ep[k]
    st = rt.store()
    f2c = torch.empty((B, 3), device=dev)
    f2h = torch.empty((B, NUM_HEADING_BIN), device=dev)
    f2s = torch.empty((B, NUM_SIZE_CLUSTER, 3), device=dev)
    call('t3d_f2', ptr(end_points[prefix + 'center']), ptr(end_points[prefix + 'heading_residuals']),
         ptr(end_points[prefix + 'size_residuals']), ptr(totals[0]), ptr(totals[2]), ptr(totals[1]), B,
         NUM_HEADING_BIN, NUM_SIZE_CLUSTER, ptr(f2c), ptr(f2h), ptr(f2s), stream())
    end_points.update({'boxpc_fit_prob': boxpc_fit_prob, 'F2_center': f2c,
                       'F2_heading_scores': end_points[prefix + 'heading_scores'], 'F2_heading_residuals': f2h,
                       'F2_size_scores': end_points[prefix + 'size_scores'], 'F2_size_residuals': f2s})
    end_points['logits'] = logits
    return logits, end_points


class Session(object):
    """Stands in for the tf.Session of test_semisup.get_model: run(fetches, feed_dict).

    Like the TF1 graph it replaces, a session is built for one static (batch_size, num_point, num_channel)
    (test_semisup.py:61-66): with cuda_graph=True the whole eval-mode forward (~40 kernel launches, launch-latency bound
    at the reference's batch of 32) is captured once into a CUDA graph on first use and replayed per run();
    feeds of any other shape take the eager launch path."""

    def __init__(self, FLAGS, store, use_oracle_mask=False, batch_size=None, num_point=None, num_channel=None, cuda_graph=True):
        self.FLAGS, self.store, self.use_oracle_mask = FLAGS, store, use_oracle_mask
        self.shape = (batch_size, num_point, num_channel)
        self.cuda_graph = bool(cuda_graph) and None not in self.shape
        self._graph = None

    def _capture(self, with_mask):
        dev = self.store.device
        B, N, C = self.shape
        self._pc = torch.zeros((B, N, C), dtype=torch.float32, device=dev)
        self._oh = torch.zeros((B, NUM_CLASS), dtype=torch.float32, device=dev)
        self._om = torch.zeros((B, N), dtype=torch.float32, device=dev) if with_mask else None
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():       # warm-up: weight packing, function attributes, allocator pools
            for _ in range(2):
                build_graph(self.FLAGS, self._pc, self._oh, oracle_mask=self._om)
        torch.cuda.current_stream(dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g), torch.no_grad():
            logits, ep = build_graph(self.FLAGS, self._pc, self._oh, oracle_mask=self._om)
        self._graph, self._ep, self._with_mask = g, ep, with_mask

    def run(self, fetches, feed_dict):
        dev = self.store.device
        T = lambda v, dt=torch.float32: torch.as_tensor(np.asarray(v)).to(device=dev, dtype=dt).contiguous() \
            if not torch.is_tensor(v) else v.to(device=dev, dtype=dt).contiguous()
        if bool(feed_dict.get('is_training_pl', False)):
            raise NotImplementedError('is_training=True')
        pc = T(feed_dict['pc_pl'])
        one_hot = T(feed_dict['one_hot_vec_pl'])
        om = T(feed_dict['y_seg_pl']) if (self.use_oracle_mask and 'y_seg_pl' in feed_dict) else None
        rt.set_default_store(self.store)
        if self.cuda_graph and tuple(pc.shape) == self.shape and rt.get_precision() == getattr(self, '_prec', rt.get_precision()):
            if self._graph is None or self._with_mask != (om is not None):
                self._prec = rt.get_precision()
                self._capture(om is not None)
            self._pc.copy_(pc)
            self._oh.copy_(one_hot)
            if om is not None:
                self._om.copy_(om)
            self._graph.replay()
            # the graph's outputs are static buffers, overwritten by the next run(): hand out copies (sess.run returns values)
            return [_clone_value(self._ep[f]) if isinstance(f, str) else f for f in fetches]
        with torch.no_grad():
            logits, ep = build_graph(self.FLAGS, pc, one_hot, oracle_mask=om)
        out = []
        for f in fetches:
            out.append(ep[f] if isinstance(f, str) else f)
        return out


def _clone_value(v):
    """Copy of an end point out of the graph's static buffers: tensors, and the tuple / dict / None end points
    ('F_pred_box_reg', 'boxpc_feats_dict', 'boxpc_fit_prob' with refine = 0) the eager path returns as they are."""
    if torch.is_tensor(v):
        return v.clone()
    if isinstance(v, (tuple, list)):
        return type(v)(_clone_value(x) for x in v)
    if isinstance(v, dict):
        return {k: _clone_value(x) for k, x in v.items()}
    return v


FLAGS = None            # the reference's module-level flags (test_semisup.py:545-548): set by the driver before get_model()
MODEL_PATH = None       # checkpoint prefix get_model() restores when no `variables` are handed in (test_semisup.py:22, :158-159)


def get_model(batch_size, num_point, num_channel, use_oracle_mask=False, FLAGS=None, variables=None, device='cuda', cuda_graph=True):
    """test_semisup.get_model (test_semisup.py:61-179; same positional signature) -> (sess, ops).  The reference reads its flags
    and the checkpoint path from module globals; so does this one when the keyword arguments are left out:
    `variables` = {TF name: array} or a VariableStore (what saver.restore would load, :158-159), default: the TF checkpoint at
    test_semisup.MODEL_PATH; `FLAGS` default: test_semisup.FLAGS."""
    flags = FLAGS if FLAGS is not None else globals()['FLAGS']
    if flags is None:
        raise ValueError('get_model: pass FLAGS= or set test_semisup.FLAGS (the reference parses them into a module global)')
    if variables is None:
        if MODEL_PATH is None:
            raise ValueError('get_model: pass variables= or set test_semisup.MODEL_PATH to a TensorFlow checkpoint prefix')
        from . import tf_checkpoint
        variables = tf_checkpoint.load_checkpoint(MODEL_PATH)
    store = variables if isinstance(variables, rt.VariableStore) else rt.VariableStore(variables, device)
    sess = Session(flags, store, use_oracle_mask, batch_size, num_point, num_channel, cuda_graph)
    ops = {k: k for k in ('pc_pl', 'one_hot_vec_pl', 'y_seg_pl', 'y_centers_pl', 'y_orient_cls_pl', 'y_orient_reg_pl',
                          'y_dims_cls_pl', 'y_dims_reg_pl', 'R0_rect_pl', 'P_pl', 'Rtilt_pl', 'K_pl', 'rot_frust_pl',
                          'box2D_pl', 'img_dim_pl', 'is_training_pl')}
    ops['logits'] = 'logits'
    ops['end_points'] = _KeyDict()
    return sess, ops


class _KeyDict(dict):
    """ops['end_points'][name] -> name, so fetch lists are written exactly like the reference's."""

    def __missing__(self, key):
        return key


def softmax(x):
    """test_semisup.py:181-185."""
    shape = x.shape
    probs = np.exp(x - np.max(x, axis=len(shape) - 1, keepdims=True))
    probs /= np.sum(probs, axis=len(shape) - 1, keepdims=True)
    return probs


def inference(sess, ops, pc, one_hot_vec, batch_size, prefix='', use_boxpc_fit_prob=False, oracle_mask=None):
    """test_semisup.py:187-260: batches through sess.run, host-side numpy scores and argmax-selects."""
    assert pc.shape[0] % batch_size == 0
    num_batches = pc.shape[0] // batch_size
    n = pc.shape[0]
    boxpc_fit_prob = np.zeros((n,))
    logits = np.zeros((n, pc.shape[1], 2))
    centers = np.zeros((n, 3))
    heading_logits = np.zeros((n, NUM_HEADING_BIN))
    heading_residuals = np.zeros((n, NUM_HEADING_BIN))
    size_logits = np.zeros((n, NUM_SIZE_CLUSTER))
    size_residuals = np.zeros((n, NUM_SIZE_CLUSTER, 3))
    scores = np.zeros((n,))
    ep = ops['end_points']
    for i in range(num_batches):
        sl = slice(i * batch_size, (i + 1) * batch_size)
        feed_dict = {ops['pc_pl']: pc[sl, ...], ops['one_hot_vec_pl']: one_hot_vec[sl, :], ops['is_training_pl']: False}
        if oracle_mask is not None:
            feed_dict.update({ops['y_seg_pl']: oracle_mask[sl]})
        run_ops = [ops['logits'], ep[prefix + 'center'], ep[prefix + 'heading_scores'], ep[prefix + 'heading_residuals'],
                   ep[prefix + 'size_scores'], ep[prefix + 'size_residuals']]
        if use_boxpc_fit_prob:
            run_ops.append(ep['boxpc_fit_prob'])
        res = [r.detach().cpu().numpy().astype(np.float64) for r in sess.run(run_ops, feed_dict=feed_dict)]
        batch_logits, batch_centers, batch_hs, batch_hr, batch_ss, batch_sr = res[:6]
        logits[sl, ...] = batch_logits
        centers[sl, ...] = batch_centers
        heading_logits[sl, ...] = batch_hs
        heading_residuals[sl, ...] = batch_hr
        size_logits[sl, ...] = batch_ss
        size_residuals[sl, ...] = batch_sr
        batch_seg_prob = softmax(batch_logits)[:, :, 1]
        batch_seg_mask = np.argmax(batch_logits, 2)
        mask_mean_prob = np.sum(batch_seg_prob * batch_seg_mask, 1)
        mask_mean_prob = mask_mean_prob / (np.sum(batch_seg_mask, 1) + 1)
        heading_prob = np.max(softmax(batch_hs), 1)
        size_prob = np.max(softmax(batch_ss), 1)
        if use_boxpc_fit_prob:
            boxpc_fit_prob[sl] = res[6]
            batch_scores = np.log(res[6] + 0.01) + np.log(mask_mean_prob + 0.01) + np.log(heading_prob + 0.01) + \
                np.log(size_prob + 0.01)
        else:
            batch_scores = np.log(mask_mean_prob + 0.01) + np.log(heading_prob + 0.01) + np.log(size_prob + 0.01)
        scores[sl] = batch_scores
    heading_cls = np.argmax(heading_logits, 1)
    size_cls = np.argmax(size_logits, 1)
    pred_seg = np.argmax(logits, 2)
    pred_orient_reg = np.array([heading_residuals[i, heading_cls[i]] for i in range(n)])
    pred_dims_reg = np.vstack([size_residuals[i, size_cls[i], :] for i in range(n)])
    return pred_seg, centers, heading_cls, pred_orient_reg, size_cls, pred_dims_reg, scores


def inference_scores(logits, heading_scores, heading_residuals, size_scores, size_residuals, fit_prob=None):
    """The numpy block of inference() (test_semisup.py:236-258) for one batch, on the device (t3d_inference_scores).
    Device tensors in -> dict of device tensors: pred_seg (B,N) uint8, mask_mean_prob, heading_cls, heading_res, size_cls,
    size_res (B,3), scores."""
    from ._lib import t3d_infer_score_args, load, check, ptr, stream
    import ctypes
    f = rt.f32
    logits, hs, hr, ss, sr = f(logits), f(heading_scores), f(heading_residuals), f(size_scores), f(size_residuals)
    fp = f(fit_prob) if fit_prob is not None else None
    B, N = logits.shape[0], logits.shape[1]
    dev = logits.device
    out = {'pred_seg': torch.empty((B, N), dtype=torch.uint8, device=dev),
           'mask_mean_prob': torch.empty((B,), dtype=torch.float32, device=dev),
           'heading_cls': torch.empty((B,), dtype=torch.int32, device=dev),
           'heading_res': torch.empty((B,), dtype=torch.float32, device=dev),
           'size_cls': torch.empty((B,), dtype=torch.int32, device=dev),
           'size_res': torch.empty((B, 3), dtype=torch.float32, device=dev),
           'scores': torch.empty((B,), dtype=torch.float32, device=dev)}
    a = t3d_infer_score_args(ptr(logits), ptr(hs), ptr(hr), ptr(ss), ptr(sr), ptr(fp), B, N, hs.shape[1], ss.shape[1],
                             ptr(out['pred_seg']), ptr(out['mask_mean_prob']), ptr(out['heading_cls']), ptr(out['heading_res']),
                             ptr(out['size_cls']), ptr(out['size_res']), ptr(out['scores']))
    check(load().t3d_inference_scores(ctypes.byref(a), stream()))
    return out


def inference_device(sess, ops, pc, one_hot_vec, batch_size, prefix='', use_boxpc_fit_prob=False, oracle_mask=None):
    """inference() with the post-processing on the device: the same 7-tuple (numpy), but per frustum only the prediction
    crosses PCIe (N bytes of pred_seg + 10 floats) instead of the raw fetches (2 N + 67 floats).  pred_seg is uint8."""
    assert pc.shape[0] % batch_size == 0
    n, ep = pc.shape[0], ops['end_points']
    parts = []
    for i in range(n // batch_size):
        sl = slice(i * batch_size, (i + 1) * batch_size)
        feed_dict = {ops['pc_pl']: pc[sl, ...], ops['one_hot_vec_pl']: one_hot_vec[sl, :], ops['is_training_pl']: False}
        if oracle_mask is not None:
            feed_dict.update({ops['y_seg_pl']: oracle_mask[sl]})
        run_ops = [ops['logits'], ep[prefix + 'center'], ep[prefix + 'heading_scores'], ep[prefix + 'heading_residuals'],
                   ep[prefix + 'size_scores'], ep[prefix + 'size_residuals']]
        if use_boxpc_fit_prob:
            run_ops.append(ep['boxpc_fit_prob'])
        res = sess.run(run_ops, feed_dict=feed_dict)
        out = inference_scores(res[0], res[2], res[3], res[4], res[5], res[6] if use_boxpc_fit_prob else None)
        out['center'] = res[1]
        parts.append(out)
    cat = lambda k: torch.cat([p[k] for p in parts], 0).cpu().numpy()
    return (cat('pred_seg'), cat('center').astype(np.float64), cat('heading_cls').astype(np.int64), cat('heading_res').astype(np.float64),
            cat('size_cls').astype(np.int64), cat('size_res').astype(np.float64), cat('scores').astype(np.float64))


def write_detection_results(result_dir, test_classes, id_list, type_list, box2d_list, center_list, heading_cls_list, heading_res_list,
                            size_cls_list, size_res_list, rot_angle_list, score_list):
    """test_semisup.py:262-294: one `<class>_pred.txt` per class, lines
    `idx cls -1 -1 -10 x1 y1 x2 y2 h w l tx ty tz ry score` (%f); the label conversion runs as one batched kernel."""
    import os
    from .roi_seg_box3d_dataset import from_prediction_to_label_format_batch
    assert result_dir is not None
    if not os.path.exists(result_dir):
        os.mkdir(result_dir)
    cls_files = {c: open(os.path.join(result_dir, c + '_pred.txt'), 'w') for c in test_classes}
    if len(center_list) > 0:
        lab = from_prediction_to_label_format_batch(np.asarray(center_list), np.asarray(heading_cls_list), np.asarray(heading_res_list),
                                                    np.asarray(size_cls_list), np.asarray(size_res_list), np.asarray(rot_angle_list))
        lab = lab.cpu().numpy().astype(np.float64)
    for i in range(len(center_list)):
        box2d, cls_name = box2d_list[i], type_list[i]
        h, w, l, tx, ty, tz, ry = lab[i]
        cls_files[cls_name].write('%d %s -1 -1 -10 %f %f %f %f %f %f %f %f %f %f %f %f\n' % (
            id_list[i], cls_name, box2d[0], box2d[1], box2d[2], box2d[3], h, w, l, tx, ty, tz, ry, score_list[i]))
    for f in cls_files.values():
        f.close()


def write_gt_results(result_dir, test_classes, test_dataset):
    """test_semisup.py:296-327: `<class>_gt.txt`, lines `idx cls -1 -1 -10 x1 y1 x2 y2 h w l tx ty tz heading`.
    test_dataset needs idx_l, box2d_l, cls_type_l, size_l, heading_l and get_box3d_center(i)."""
    import os
    assert result_dir is not None
    if not os.path.exists(result_dir):
        os.mkdir(result_dir)
    cls_files = {c: open(os.path.join(result_dir, c + '_gt.txt'), 'w') for c in test_classes}
    for i in range(len(test_dataset)):
        box2d, cls_name = test_dataset.box2d_l[i], test_dataset.cls_type_l[i]
        l, w, h = test_dataset.size_l[i]
        tx, ty, tz = test_dataset.get_box3d_center(i)
        cls_files[cls_name].write('%d %s -1 -1 -10 %f %f %f %f %f %f %f %f %f %f %f\n' % (
            test_dataset.idx_l[i], cls_name, box2d[0], box2d[1], box2d[2], box2d[3], h, w, l, tx, ty, tz, test_dataset.heading_l[i]))
    for f in cls_files.values():
        f.close()


def fill_files(output_dir, to_fill_filename_list):
    """test_semisup.py:329-334."""
    import os
    for filename in to_fill_filename_list:
        filepath = os.path.join(output_dir, filename)
        if not os.path.exists(filepath):
            open(filepath, 'w').close()


def _pad_batch(t, batch_size):
    """The reference feeds a fixed-size batch (static TF graph) and leaves the unused tail rows zero (:470-471)."""
    if t.shape[0] == batch_size:
        return t
    out = torch.zeros((batch_size,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    out[:t.shape[0]] = t
    return out


def main_batch(test_dataset, test_classes, num_class, num_point, num_channel, prefix='', semi_type=None, use_boxpc_fit_prob=False,
               sess_ops=None, output_filename=None, result_dir=None, verbose=False, batch_size=32, FLAGS=None, variables=None):
    """test_semisup.py:430-520: batches of 32 through get_batch -> inference -> the 14-list `predictions`
    [ps, seg_gt, seg_pred, center, heading_cls, heading_res, size_cls, size_res, rot_angle, score, cls, file_num, box2d, box3d],
    optional gz pickle and per-class result files.  test_dataset is roi_seg_box3d_dataset.ROISegBoxDataset (device batches);
    the post-processing runs on the device (inference_device).  sess_ops=None builds the session from (FLAGS, variables)."""
    from .utils import save_zipped_pickle
    lists = [[] for _ in range(14)]
    test_idxs = np.arange(0, len(test_dataset))
    num_batches = int((len(test_dataset) + batch_size - 1) / batch_size)
    sess, ops = sess_ops if sess_ops is not None else get_model(batch_size, num_point, num_channel, FLAGS=FLAGS, variables=variables)
    idx, iou_sum = 0, 0.0
    for batch_idx in range(num_batches):
        start_idx, end_idx = batch_idx * batch_size, min(len(test_dataset), (batch_idx + 1) * batch_size)
        cur = end_idx - start_idx
        b = test_dataset.get_batch(test_idxs, start_idx, end_idx, num_point, num_channel)
        batch_data, batch_label, batch_rot_angle, batch_one_hot_vec = b[0], b[2], b[11], b[13]
        out = inference_device(sess, ops, _pad_batch(batch_data, batch_size), _pad_batch(batch_one_hot_vec, batch_size), batch_size,
                               prefix=prefix, use_boxpc_fit_prob=use_boxpc_fit_prob)
        batch_output, center_pred, hclass_pred, hres_pred, sclass_pred, sres_pred, scores = out
        data_h, label_h, rot_h, oh_h = (t.cpu().numpy() for t in (batch_data, batch_label, batch_rot_angle, batch_one_hot_vec))
        for i in range(cur):                                  # segmentation IoU on the un-duplicated points (:479-490)
            _, unique_idx = np.unique(data_h[i], axis=0, return_index=True)
            y_pred, y_true = batch_output[i][unique_idx].astype(np.int64), label_h[i][unique_idx].astype(np.int64)
            iou_sum += float(np.sum(y_pred & y_true)) / (np.sum(y_pred | y_true) + 1)
        for i in range(cur):
            vals = (data_h[i], label_h[i], batch_output[i], center_pred[i], hclass_pred[i], hres_pred[i], sclass_pred[i], sres_pred[i],
                    rot_h[i], scores[i], int(np.argmax(oh_h[i])), test_dataset.idx_l[idx], test_dataset.box2d_l[idx],
                    test_dataset.box3d_l[idx])
            for l, v in zip(lists, vals):
                l.append(v)
            idx += 1
    if verbose:
        print('Mean segmentation IOU: %f' % (iou_sum / max(len(test_dataset.idx_l), 1)))
    predictions = lists
    if output_filename is not None:
        save_zipped_pickle(predictions, output_filename)
    if result_dir is not None:
        write_detection_results(result_dir, test_classes, test_dataset.idx_l, test_dataset.cls_type_l, test_dataset.box2d_l,
                                lists[3], lists[4], lists[5], lists[6], lists[7], lists[8], lists[9])
    return predictions


def main_batch_from_rgb_detection(test_dataset, test_classes, num_class, num_point, num_channel, prefix='', semi_type=None,
                                  use_boxpc_fit_prob=False, use_oracle_mask=False, sess_ops=None, output_filename=None, result_dir=None,
                                  verbose=False, batch_size=32, FLAGS=None, variables=None):
    """test_semisup.py:336-428: the 7-list (rgb detection) flow; the score kept is the 2D detector's probability (:401) and
    predictions = [ps, None, seg_pred, center, heading_cls, heading_res, size_cls, size_res, rot_angle, score, cls, file_num,
    box2d, None]."""
    from .utils import save_zipped_pickle
    lists = [[] for _ in range(14)]
    lists[1] = lists[13] = None
    test_idxs = np.arange(0, len(test_dataset))
    num_batches = int((len(test_dataset) + batch_size - 1) / batch_size)
    sess, ops = sess_ops if sess_ops is not None else get_model(batch_size, num_point, num_channel, use_oracle_mask=use_oracle_mask,
                                                                 FLAGS=FLAGS, variables=variables)
    idx = 0
    for batch_idx in range(num_batches):
        start_idx, end_idx = batch_idx * batch_size, min(len(test_dataset), (batch_idx + 1) * batch_size)
        cur = end_idx - start_idx
        batch_data, _, batch_rot_angle, batch_rgb_prob, batch_one_hot_vec, batch_oracle_y_seg = \
            test_dataset.get_batch(test_idxs, start_idx, end_idx, num_point, num_channel, from_rgb_detection=True)
        om = _pad_batch(batch_oracle_y_seg, batch_size) if use_oracle_mask else None
        out = inference_device(sess, ops, _pad_batch(batch_data, batch_size), _pad_batch(batch_one_hot_vec, batch_size), batch_size,
                               prefix=prefix, use_boxpc_fit_prob=use_boxpc_fit_prob, oracle_mask=om)
        batch_output, center_pred, hclass_pred, hres_pred, sclass_pred, sres_pred, _ = out
        data_h, rot_h, prob_h, oh_h = (t.cpu().numpy() for t in (batch_data, batch_rot_angle, batch_rgb_prob, batch_one_hot_vec))
        for i in range(cur):
            vals = {0: data_h[i], 2: batch_output[i], 3: center_pred[i], 4: hclass_pred[i], 5: hres_pred[i], 6: sclass_pred[i],
                    7: sres_pred[i], 8: rot_h[i], 9: prob_h[i], 10: int(np.argmax(oh_h[i])), 11: test_dataset.idx_l[idx],
                    12: test_dataset.box2d_l[idx]}
            for k, v in vals.items():
                lists[k].append(v)
            idx += 1
    predictions = lists
    if output_filename is not None:
        save_zipped_pickle(predictions, output_filename)
    if result_dir is not None:
        write_detection_results(result_dir, test_classes, test_dataset.idx_l, test_dataset.cls_type_l, test_dataset.box2d_l,
                                lists[3], lists[4], lists[5], lists[6], lists[7], lists[8], lists[9])
    return predictions
