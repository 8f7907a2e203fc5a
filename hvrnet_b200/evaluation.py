"""Results format and VID mAP evaluation - the step after the hot path.

``eval_map`` restates what tools/vid_eval.py:11-52 asks of mmdet/core/evaluation/mean_ap.py
(:376-438 tpfp_default, :441-456 get_cls_results, :475-585 eval_map, :9-53 average_precision
'area' mode, bbox_overlaps.py with the +1 pixel convention): per class, detections of all frames
are matched greedily in score order to the best-overlapping ground truth (IoU >= 0.5, each gt
once, ignored gts neither tp nor fp), AP is the area under the monotone precision envelope, mAP
the mean over classes that have ground truth.  Host-side numpy, as in the reference.

``dump_results`` / ``load_results`` write the per-frame ``bbox_result`` lists (30 float32 [k,5]
arrays per frame) the way the reference's drivers hand them to mmcv.dump (tools/hnl_test.py:770-800).
"""
import pickle

import numpy as np


def bbox_overlaps(b1, b2):
    """IoU matrix [n,k] with the +1 pixel convention, float32 (bbox_overlaps.py:4-49)."""
    b1 = b1.astype(np.float32)
    b2 = b2.astype(np.float32)
    if b1.shape[0] * b2.shape[0] == 0:
        return np.zeros((b1.shape[0], b2.shape[0]), dtype=np.float32)
    a1 = (b1[:, 2] - b1[:, 0] + 1) * (b1[:, 3] - b1[:, 1] + 1)
    a2 = (b2[:, 2] - b2[:, 0] + 1) * (b2[:, 3] - b2[:, 1] + 1)
    xs = np.maximum(b1[:, None, 0], b2[None, :, 0])
    ys = np.maximum(b1[:, None, 1], b2[None, :, 1])
    xe = np.minimum(b1[:, None, 2], b2[None, :, 2])
    ye = np.minimum(b1[:, None, 3], b2[None, :, 3])
    ov = np.maximum(xe - xs + 1, 0) * np.maximum(ye - ys + 1, 0)
    return (ov / (a1[:, None] + a2[None, :] - ov)).astype(np.float32)


def tpfp(dets, gts, gt_ignore, iou_thr):
    """tp / fp flags (float32 [n]) of one image's detections of one class (mean_ap.py:376-438)."""
    n = dets.shape[0]
    tp = np.zeros(n, dtype=np.float32)
    fp = np.zeros(n, dtype=np.float32)
    if gts.shape[0] == 0:
        fp[:] = 1
        return tp, fp
    ious = bbox_overlaps(dets[:, :4], gts[:, :4])
    best, arg = ious.max(axis=1), ious.argmax(axis=1)
    covered = np.zeros(gts.shape[0], dtype=bool)
    for i in np.argsort(-dets[:, -1]):
        if best[i] >= iou_thr:
            j = arg[i]
            if not gt_ignore[j]:
                if not covered[j]:
                    covered[j] = True
                    tp[i] = 1
                else:
                    fp[i] = 1
        else:
            fp[i] = 1
    return tp, fp


def average_precision(recalls, precisions):
    """Area under the monotone precision envelope (mean_ap.py:9-53, mode 'area')."""
    mrec = np.concatenate([[0.0], recalls, [1.0]]).astype(recalls.dtype)
    mpre = np.concatenate([[0.0], precisions, [0.0]]).astype(precisions.dtype)
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    ind = np.where(mrec[1:] != mrec[:-1])[0]
    return np.float32(np.sum((mrec[ind + 1] - mrec[ind]) * mpre[ind + 1]))


def eval_map(det_results, gt_bboxes, gt_labels, gt_ignore=None, iou_thr=0.5):
    """det_results: list over frames of lists over classes of [k,5] arrays; gt_labels 1-based.
    Returns (mAP, [dict(num_gts, num_dets, recall, precision, ap) per class])."""
    assert len(det_results) == len(gt_bboxes) == len(gt_labels)
    num_classes = len(det_results[0])
    gt_labels = [l if l.ndim == 1 else l[:, 0] for l in gt_labels]
    results = []
    eps = np.finfo(np.float32).eps
    for c in range(num_classes):
        tps, fps, dets_all, num_gts = [], [], [], 0
        for j, det in enumerate(det_results):
            sel = gt_labels[j] == c + 1
            g = gt_bboxes[j][sel] if gt_bboxes[j].shape[0] > 0 else gt_bboxes[j]
            ign = np.zeros(g.shape[0], dtype=bool) if gt_ignore is None else np.asarray(gt_ignore[j])[sel].astype(bool)
            t, f = tpfp(det[c], g, ign, iou_thr)
            tps.append(t)
            fps.append(f)
            dets_all.append(det[c])
            num_gts += int(np.sum(~ign))
        dets_all = np.vstack(dets_all)
        order = np.argsort(-dets_all[:, -1])
        tp = np.cumsum(np.hstack(tps)[order])
        fp = np.cumsum(np.hstack(fps)[order])
        rec = tp / np.maximum(num_gts, eps)
        prec = tp / np.maximum(tp + fp, eps)
        results.append(dict(num_gts=num_gts, num_dets=dets_all.shape[0], recall=rec, precision=prec,
                            ap=average_precision(rec, prec)))
    aps = [r['ap'] for r in results if r['num_gts'] > 0]
    return (float(np.array(aps).mean()) if aps else 0.0), results


def dump_results(results, path):
    with open(path, 'wb') as f:
        pickle.dump(results, f, protocol=2)       # mmcv.dump(..., 'x.pkl') default protocol


def load_results(path):
    with open(path, 'rb') as f:
        return pickle.load(f)
