"""Inference graph + runner: restatement of sunrgbd_detection/test_semisup.py:61-149
(get_model: model F + BoxPC refine loop -> F2_* end points) and :181-260 (inference: batches,
numpy softmax / argmax / log-score).
"""
import numpy as np
import torch

from . import tf_util, semisup_v1_sunrgbd as MODEL, boxpc_sunrgbd
from transferable3d_b200.constants import NUM_HEADING_BIN, NUM_SIZE_CLUSTER


def run_graph(vs, FLAGS, pc, one_hot_vec, box2D=None, img_dim=None, oracle_mask=None, is_training=False):
    """The body of test_semisup.get_model (test_semisup.py:74-149) evaluated on one batch."""
    norm_box2D = None
    if box2D is not None and img_dim is not None:
        norm_box2D = tf_util.tf_normalize_2D_bboxes(box2D, img_dim)
    pred, end_points = MODEL.get_semi_model(pc, None, None, one_hot_vec, is_training, FLAGS.use_one_hot, vs,
                                            oracle_mask=oracle_mask, norm_box2D=norm_box2D, c=FLAGS)
    logits = pred[0]
    prefix = 'F_'
    n_refine = int(FLAGS.refine)
    curr_box = end_points[prefix + 'pred_box_reg']
    curr_center_reg, curr_size_reg, curr_angle_reg = curr_box
    boxpc_fit_prob = None
    total_delta_center = torch.zeros_like(curr_center_reg)
    total_delta_angle = torch.zeros_like(curr_angle_reg)
    total_delta_size = torch.zeros_like(curr_size_reg)
    for i in range(n_refine):
        if FLAGS.mask_pc_for_boxpc:
            mask = torch.argmax(logits, dim=2).to(pc.dtype).unsqueeze(2)
            fake_box_pc = (curr_box, pc * mask)
        else:
            fake_box_pc = (curr_box, pc)
        with vs.variable_scope('D_boxpc_branch'):
            _, ep = boxpc_sunrgbd.get_model(fake_box_pc, False, one_hot_vec, vs, use_one_hot_vec=False, c=FLAGS)
        boxpc_fit_prob = torch.softmax(ep['boxpc_fit_logits'], dim=1)[:, 1]
        # test_semisup.py:86 sets FLAGS.SEMI_WEIGH_BOXPC_DELTA_DURING_TEST = False before building the graph, so the
        # `1 - logits_for_weigh` branch (:117-119) is never taken at test time
        weight = torch.ones_like(ep['logits_for_weigh'])
        delta_center = ep['boxpc_delta_center'] * weight.unsqueeze(1)
        delta_angle = ep['boxpc_delta_angle'] * weight
        delta_size = ep['boxpc_delta_size'] * weight.unsqueeze(1)
        c0, s0, a0 = curr_box
        curr_box = (c0 - delta_center, s0 - delta_size, a0 - delta_angle)
        total_delta_center = total_delta_center + delta_center
        total_delta_angle = total_delta_angle + delta_angle
        total_delta_size = total_delta_size + delta_size
        end_points['boxpc_delta_center'] = ep['boxpc_delta_center']
        end_points['boxpc_delta_size'] = ep['boxpc_delta_size']
        end_points['boxpc_delta_angle'] = ep['boxpc_delta_angle']
        end_points['boxpc_feats_dict'] = ep['boxpc_feats_dict']
        end_points['pred_boxpc_fit'] = ep['pred_boxpc_fit']
    end_points.update({
        'boxpc_fit_prob': boxpc_fit_prob,
        'F2_center': end_points[prefix + 'center'] - total_delta_center,
        'F2_heading_scores': end_points[prefix + 'heading_scores'],
        'F2_heading_residuals': end_points[prefix + 'heading_residuals'] - total_delta_angle.unsqueeze(1).repeat(1, 12),
        'F2_size_scores': end_points[prefix + 'size_scores'],
        'F2_size_residuals': end_points[prefix + 'size_residuals'] - total_delta_size.unsqueeze(1).repeat(1, 10, 1)})
    end_points['logits'] = logits
    return logits, end_points


def softmax(x):
    """test_semisup.py:181-185."""
    shape = x.shape
    probs = np.exp(x - np.max(x, axis=len(shape) - 1, keepdims=True))
    probs /= np.sum(probs, axis=len(shape) - 1, keepdims=True)
    return probs


def inference(vs, FLAGS, pc, one_hot_vec, batch_size, prefix='', use_boxpc_fit_prob=False):
    """test_semisup.py:187-260; host post-processing is float64 numpy as in the reference."""
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
    for i in range(num_batches):
        sl = slice(i * batch_size, (i + 1) * batch_size)
        with torch.no_grad():
            _, ep = run_graph(vs, FLAGS, torch.as_tensor(pc[sl]).to(vs.dtype),
                              torch.as_tensor(one_hot_vec[sl]).to(vs.dtype))
        b_logits = ep['logits'].numpy().astype(np.float64)
        logits[sl] = b_logits
        centers[sl] = ep[prefix + 'center'].numpy()
        b_hs = ep[prefix + 'heading_scores'].numpy().astype(np.float64)
        heading_logits[sl] = b_hs
        heading_residuals[sl] = ep[prefix + 'heading_residuals'].numpy()
        b_ss = ep[prefix + 'size_scores'].numpy().astype(np.float64)
        size_logits[sl] = b_ss
        size_residuals[sl] = ep[prefix + 'size_residuals'].numpy()
        seg_prob = softmax(b_logits)[:, :, 1]
        seg_mask = np.argmax(b_logits, 2)
        mask_mean_prob = np.sum(seg_prob * seg_mask, 1) / (np.sum(seg_mask, 1) + 1)
        heading_prob = np.max(softmax(b_hs), 1)
        size_prob = np.max(softmax(b_ss), 1)
        if use_boxpc_fit_prob:
            fp = ep['boxpc_fit_prob'].numpy().astype(np.float64)
            boxpc_fit_prob[sl] = fp
            b_scores = np.log(fp + 0.01) + np.log(mask_mean_prob + 0.01) + np.log(heading_prob + 0.01) + np.log(size_prob + 0.01)
        else:
            b_scores = np.log(mask_mean_prob + 0.01) + np.log(heading_prob + 0.01) + np.log(size_prob + 0.01)
        scores[sl] = b_scores
    heading_cls = np.argmax(heading_logits, 1)
    size_cls = np.argmax(size_logits, 1)
    pred_seg = np.argmax(logits, 2)
    pred_orient_reg = np.array([heading_residuals[i, heading_cls[i]] for i in range(n)])
    pred_dims_reg = np.vstack([size_residuals[i, size_cls[i], :] for i in range(n)])
    return pred_seg, centers, heading_cls, pred_orient_reg, size_cls, pred_dims_reg, scores


def inference_scores(logits, heading_scores, heading_residuals, size_scores, size_residuals, fit_prob=None):
    """The per-batch numpy block of test_semisup.inference (test_semisup.py:236-258) on given fetches (float64):
    -> pred_seg, mask_mean_prob, heading_cls, heading_res, size_cls, size_res, scores."""
    logits = np.asarray(logits, dtype=np.float64)
    hs, ss = np.asarray(heading_scores, dtype=np.float64), np.asarray(size_scores, dtype=np.float64)
    seg_prob = softmax(logits)[:, :, 1]
    seg_mask = np.argmax(logits, 2)
    mask_mean_prob = np.sum(seg_prob * seg_mask, 1) / (np.sum(seg_mask, 1) + 1)
    heading_prob = np.max(softmax(hs), 1)
    size_prob = np.max(softmax(ss), 1)
    scores = np.log(mask_mean_prob + 0.01) + np.log(heading_prob + 0.01) + np.log(size_prob + 0.01)
    if fit_prob is not None:
        scores = scores + np.log(np.asarray(fit_prob, dtype=np.float64) + 0.01)
    heading_cls, size_cls = np.argmax(hs, 1), np.argmax(ss, 1)
    n = logits.shape[0]
    heading_res = np.array([heading_residuals[i, heading_cls[i]] for i in range(n)])
    size_res = np.vstack([size_residuals[i, size_cls[i], :] for i in range(n)])
    return seg_mask, mask_mean_prob, heading_cls, heading_res, size_cls, size_res, scores
