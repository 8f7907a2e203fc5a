"""Streaming window scheduler with per-frame caches.

The reference's test loop (tools/hnl_test.py:359-463) keeps a deque of C4 maps and, for EVERY
key frame, recomputes C5, the RPN, the proposals, RoIAlign and fc_new_1 for ALL T frames of the
window (hnmb_rcnn.py:202-203,596-599): 2020 GFLOP per key frame at T=15, of which only ~570 are
new work.  Everything up to and including fc_new_1 is a per-frame function, so this scheduler
computes it once when a frame arrives and caches, per frame,

    proposals [P,5]      fc_new_1 rows f1 [P,1024] (split)      f1^T [1024,P] (split)

and for each key frame runs only the window-dependent part (relation stages 1-4, branch heads,
decode + NMS) on the concatenated cached rows.  The kernels and their per-element arithmetic are
the ones of the as-executed path, so the detections are bit-identical to
``model(x=window, forward_feat=True)`` (tests/test_gpu_pipeline.py::test_streaming_bit_identical).

This is a scheduling extension, NOT the headline bench path: bench.py's `value` / `e2e` time the
path as the reference executes it; the streaming figure is reported separately and labelled.
"""
from collections import deque

import torch

from . import engine, ops
from .models import bbox2result


class _Frame:
    __slots__ = ('props', 'count', 'f1', 'f1T', 'meta')


class StreamingDetector:
    """One video stream.  ``push(img, img_meta)`` -> detections of the window's key frame once the
    window is full (list over head outputs of per-class [k,5] arrays), else None."""

    def __init__(self, model, window=None):
        self.m = model
        self.T = int(window or model.bbox_head.t_dim)
        self.frames = deque(maxlen=self.T)

    # ---- per-frame stage: everything that does not depend on the window ----------------------
    def _frame_stage(self, img, meta):
        m = self.m
        c4 = m.backbone.forward_split(img)                                  # [B,h,w,1024]
        c5 = m.shared_head.forward_nhwc(c4)
        props, counts = m.rpn_head.get_proposals(c4, meta['img_shape'], m.test_cfg.rpn)
        B, P = props.shape[0], props.shape[1]
        cnt = counts.cpu().tolist()
        out = []
        packed = m.bbox_head.packed(img.device)
        for b in range(B):
            n = cnt[b]
            rois = torch.cat([props.new_full((n, 1), float(b)), props[b, :n, :4]], -1).contiguous()
            rows = m.bbox_roi_extractor.roi_layers[0].forward_nhwc_split(c5, rois)
            f1, f1T = engine.head_fc1(packed, rows)
            fr = _Frame()
            fr.props, fr.count, fr.f1, fr.f1T, fr.meta = props[b, :n], n, f1, f1T, meta
            out.append(fr)
        return out

    def push(self, img, img_meta):
        if not img.is_cuda:
            img = img.cuda(non_blocking=True)
        for fr in self._frame_stage(img, img_meta):
            self.frames.append(fr)
        if len(self.frames) < self.T:
            return None
        return self.detect()

    # ---- window stage ---------------------------------------------------------------------------
    def detect(self, rescale=True):
        m = self.m
        frs = list(self.frames)
        key = m.key_dim
        start = sum(f.count for f in frs[:key])
        length = frs[key].count
        N = sum(f.count for f in frs)
        dev = frs[0].f1.hi.device
        f1 = ops.Split(torch.cat([f.f1.hi[:f.count] for f in frs], 0), torch.cat([f.f1.lo[:f.count] for f in frs], 0))
        f1T = ops.Split.zeros((f1.shape[1], ops.round_up(N, 64)), dev)
        torch.cat([f.f1T.hi[:, :f.count] for f in frs], 1, out=f1T.hi[:, :N])
        torch.cat([f.f1T.lo[:, :f.count] for f in frs], 1, out=f1T.lo[:, :N])
        head = m.bbox_head
        packed = head.packed(dev)
        if head.kind == 'hrnmp':
            o1, o2, _ = engine.hrnmp_forward_test(packed, None, start, length, f1=f1, f1T=f1T)
            outs = [o1, o2]
        else:
            outs = [engine.selsa_forward(packed, None, start, length, f1=f1, f1T=f1T)]
        kf = frs[key]
        rois_key = torch.cat([kf.props.new_zeros((length, 1)), kf.props[:, :4]], -1).contiguous()
        meta = frs[0].meta                                                   # frame 0's meta, hnmb_rcnn.py:603-604
        res = []
        for o in outs:
            cls, reg = head._split_out(o)
            d, l, k = head.get_det_bboxes(rois_key, cls, reg, meta['img_shape'], meta['scale_factor'], rescale=rescale,
                                          cfg=m.test_cfg.rcnn)
            kk = int(k.item())
            res.append(bbox2result(d[:kk], l[:kk], head.num_classes))
        return res
