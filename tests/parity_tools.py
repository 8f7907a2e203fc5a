"""End-to-end INDEX parity of the CUDA path against the CPU oracle (test infrastructure).

north_star: "bit-exact proposal/NMS indices, bbox/cls tensors within 1e-3 rel fp32".  Index decisions are
discontinuous functions of float tensors that agree only to the arithmetic's error (measured ~1e-4 of the
tensor's range), so free-running from frames two kinds of statements can be made, and this module makes both:

 (A) REPLAY (exact): the oracle's index logic run on the DEVICE's own float tensors (RPN maps copied to the
     host; head outputs copied to the host) must return the device's indices bit for bit - anchor index of
     every proposal, in order, and (label, roi index) of every detection, in order.  Any mismatch here is a bug
     in the CUDA index logic (sort order, tie rule, IoU arithmetic, suppression, top-k), at full size.
 (B) FREE-RUNNING (quantified): the oracle end to end on its own tensors against the device end to end.
     Mismatches are allowed only as NEAR-TIES: for the first position at which the two kept lists of a frame
     diverge, some decision that separates them (a logit order, an IoU against the NMS threshold, the top-k
     cut-off, a class score against score_thr) must have a margin below the measured arithmetic error of the
     quantity it compares.  Everything after the first divergence is a legitimate cascade of a greedy,
     sequential algorithm and is reported as a rate, not analysed.

Used by tests/parity_report.py (prints the numbers quoted in DESIGN.md / profiles/) and by
tests/test_gpu_pipeline.py (asserts them)."""
import numpy as np
import torch

from oracle import cref, ref_torch as R

RPN_CFG = dict(nms_pre=6000, nms_post=300, max_num=300, nms_thr=0.7)


def anchors_for(h, w, stride=16, scales=(4, 8, 16, 32), ratios=(0.5, 1.0, 2.0)):
    return R.grid_anchors(R.gen_base_anchors(stride, scales, ratios), (h, w), stride)


def proposal_trace(cls, reg, anchors, img_shape, cfg=RPN_CFG):
    """rpn_head.py:55-104 as oracle/ref_torch.rpn_proposals_single evaluates it (same functions, the C NMS
    instead of the Python loop), keeping the intermediate index lists.  cls (A,h,w), reg (4A,h,w) CPU fp32."""
    logits = cls.permute(1, 2, 0).reshape(-1)
    deltas = reg.permute(1, 2, 0).reshape(-1, 4)
    top = R.argsort_desc_stable(logits)[:cfg['nms_pre']] if logits.shape[0] > cfg['nms_pre'] else torch.arange(logits.shape[0])
    boxes = R.delta2bbox(anchors[top], deltas[top], max_shape=img_shape)
    keep = cref.nms(torch.cat([boxes, logits[top][:, None]], -1), cfg['nms_thr'], strict_gt=True)
    keep = keep[R.argsort_desc_stable(logits[top][keep])]
    keep = keep[:cfg['nms_post']][:cfg['max_num']]
    scores = logits[top].sigmoid()
    return dict(logits=logits, top=top, boxes=boxes, keep=keep, anchor=top[keep],
                props=torch.cat([boxes[keep], scores[keep][:, None]], -1))


def device_maps_to_oracle_layout(maps, A=12):
    """engine.rpn_forward output [T,h,w,64] (columns [0,A) logits, [A,5A) deltas a*4+d) -> (cls [T,A,h,w],
    reg [T,4A,h,w]) CPU, the layout rpn_head.py:30-35 returns."""
    m = maps.detach().float().cpu()
    return m[..., :A].permute(0, 3, 1, 2).contiguous(), m[..., A:5 * A].permute(0, 3, 1, 2).contiguous()


def first_divergence(a, b):
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return i
    return None if len(a) == len(b) else n


def _iou(b1, b2):
    return float(R._iou_row(b1, b2[None])[0])


def explain_proposal_divergence(to, td, thr=0.7):
    """to / td: proposal_trace of one frame on the oracle's maps / on the device's maps.  Returns None when the
    kept anchor lists are identical, else a dict describing the FIRST divergence: the smallest margin among the
    decisions that can separate the two lists there, next to the measured arithmetic error of the compared
    quantity (logit: |oracle - device| at the anchors involved; IoU: |IoU on oracle boxes - IoU on device
    boxes| for the pair involved)."""
    ao, ad = to['anchor'].tolist(), td['anchor'].tolist()
    j = first_divergence(ao, ad)
    if j is None:
        return None
    lo, ld = to['logits'], td['logits']
    tol_l = 1e-3 * float(lo.abs().max())     # north_star tolerance of the float tensor the decision compares
    pos_o = {int(a): i for i, a in enumerate(to['top'].tolist())}
    pos_d = {int(a): i for i, a in enumerate(td['top'].tolist())}
    cands = []          # (margin, error bound of the compared quantity, description)
    xs = [a for a in (ao[j] if j < len(ao) else None, ad[j] if j < len(ad) else None) if a is not None]
    if len(xs) == 2:
        x, y = xs
        gap = abs(float(lo[x] - lo[y]))
        err = abs(float(lo[x] - ld[x])) + abs(float(lo[y] - ld[y]))
        cands.append((gap, err, tol_l, 'logit order of anchors %d / %d' % (x, y)))
    for z in xs:
        # cut-off of the pre-NMS top-k: z against the last anchor that made it
        for t, l in ((to, lo), (td, ld)):
            last = int(t['top'][-1])
            cands.append((abs(float(l[z] - l[last])), abs(float(lo[z] - ld[z])) + abs(float(lo[last] - ld[last])), tol_l,
                          'top-%d cut-off for anchor %d' % (len(t['top']), z)))
        # suppression of z by a box kept before position j (identical prefix): IoU against the threshold
        if z in pos_o and z in pos_d:
            bo, bd = to['boxes'][pos_o[z]], td['boxes'][pos_d[z]]
            for k in ao[:j]:
                io, idv = _iou(to['boxes'][pos_o[k]], bo), _iou(td['boxes'][pos_d[k]], bd)
                if (io > thr) != (idv > thr):
                    cands.append((min(abs(io - thr), abs(idv - thr)), abs(io - idv), 1e-3,
                                  'IoU(%d, %d) = %.7f / %.7f against %.2f' % (k, z, io, idv, thr)))
    # a divergence caused by an earlier, invisible difference: a box that is suppressed on one side only
    # because the ORDER of two nearly tied logits among the candidates differs (the lower-ranked of an
    # overlapping pair is the one suppressed)
    if not any(m <= e * 4 + 1e-7 and e <= t_ for m, e, t_, _ in cands):
        n = min(len(to['top']), len(td['top']))
        d = torch.nonzero(to['top'][:n] != td['top'][:n]).reshape(-1)
        for i in d[:2000].tolist():
            x, y = int(to['top'][i]), int(td['top'][i])
            cands.append((abs(float(lo[x] - lo[y])), abs(float(lo[x] - ld[x])) + abs(float(lo[y] - ld[y])), tol_l,
                          'candidate order of anchors %d / %d at rank %d' % (x, y, i)))
    return _verdict(j, cands)


def _verdict(j, cands):
    """cands: (margin, measured error of the compared quantity, tolerance of that quantity, description).  A
    near-tie = the margin lies inside (4x) the measured error AND that error is itself within the tolerance."""
    ok = [c for c in cands if c[0] <= 4 * c[1] + 1e-7 and c[1] <= c[2]]
    m, e, t, what = min(ok, key=lambda c: c[0]) if ok else min(cands, key=lambda c: c[0] - 4 * c[1])
    return dict(position=j, margin=m, error=e, tolerance=t, what=what, near_tie=bool(ok))


def det_trace(rois, cls, reg, img_shape, scale_factor=1.0, rescale=True, score_thr=0.001, iou_thr=0.3, max_num=300):
    """get_det_bboxes + multiclass_nms (hrnmp_bbox_head.py:1009-1052, bbox_nms.py:6-66) as the oracle evaluates
    them, returning for every detection (label, roi index) in output order plus the decoded boxes / scores."""
    boxes, scores = R.decode_scores_boxes(rois, cls, reg, img_shape, scale_factor, rescale)
    dets, labels, rows = [], [], []
    for c in range(1, scores.shape[1]):
        m = scores[:, c] > score_thr
        if not m.any():
            continue
        idx = torch.nonzero(m).reshape(-1)
        d = torch.cat([boxes[m], scores[m, c][:, None]], 1)
        keep = cref.nms(d, iou_thr, strict_gt=True)
        dets.append(d[keep])
        labels.append(torch.full((keep.shape[0],), c - 1, dtype=torch.long))
        rows.append(idx[keep])
    if not dets:
        return dict(dets=boxes.new_zeros((0, 5)), labels=torch.zeros(0, dtype=torch.long),
                    rows=torch.zeros(0, dtype=torch.long), boxes=boxes, scores=scores)
    dets, labels, rows = torch.cat(dets), torch.cat(labels), torch.cat(rows)
    if dets.shape[0] > max_num:
        order = R.argsort_desc_stable(dets[:, 4])[:max_num]
        dets, labels, rows = dets[order], labels[order], rows[order]
    return dict(dets=dets, labels=labels, rows=rows, boxes=boxes, scores=scores)


def explain_det_divergence(do, dd, score_thr=0.001, iou_thr=0.3, max_num=300):
    """do / dd: det_trace on the oracle's head outputs / on the device's head outputs (same rois).  None when the
    (label, roi) lists are identical, else the first divergence with the smallest separating margin."""
    lo = list(zip(do['labels'].tolist(), do['rows'].tolist()))
    ld = list(zip(dd['labels'].tolist(), dd['rows'].tolist()))
    j = first_divergence(lo, ld)
    if j is None:
        return None
    so, sd = do['scores'], dd['scores']
    cands = []
    items = [p for p in (lo[j] if j < len(lo) else None, ld[j] if j < len(ld) else None) if p is not None]
    for (lab, row) in items:
        c = lab + 1
        err = abs(float(so[row, c] - sd[row, c]))
        cands.append((min(abs(float(so[row, c]) - score_thr), abs(float(sd[row, c]) - score_thr)), err, 1e-3,
                      'score of roi %d class %d against score_thr' % (row, lab)))
        # suppression inside the class: IoU against every other roi of the class that is above threshold
        for other in torch.nonzero((so[:, c] > score_thr) | (sd[:, c] > score_thr)).reshape(-1).tolist():
            if other == row:
                continue
            io = _iou(do['boxes'][other], do['boxes'][row])
            idv = _iou(dd['boxes'][other], dd['boxes'][row])
            if min(abs(io - iou_thr), abs(idv - iou_thr)) < 1e-3:
                cands.append((min(abs(io - iou_thr), abs(idv - iou_thr)), abs(io - idv) + 1e-7, 1e-3,
                              'IoU(roi %d, roi %d) class %d against iou_thr' % (other, row, lab)))
            gap = abs(float(so[row, c] - so[other, c]))
            if gap < 1e-3:
                cands.append((gap, err + abs(float(so[other, c] - sd[other, c])), 1e-3,
                              'score order of rois %d / %d class %d' % (row, other, lab)))
    if len(items) == 2 and (len(lo) >= max_num or len(ld) >= max_num):
        # more than max_num survivors: the output is ordered by score over all classes (bbox_nms.py:57-61), so two
        # detections of any classes with nearly equal scores may swap
        (l1, r1), (l2, r2) = items
        cands.append((abs(float(so[r1, l1 + 1] - so[r2, l2 + 1])),
                      abs(float(so[r1, l1 + 1] - sd[r1, l1 + 1])) + abs(float(so[r2, l2 + 1] - sd[r2, l2 + 1])), 1e-3,
                      'score order of (class %d, roi %d) / (class %d, roi %d) in the top-%d' % (l1, r1, l2, r2, max_num)))
    if len(lo) > max_num - 1 or len(ld) > max_num - 1:
        # top-k cut by score over all classes
        for (lab, row) in items:
            for d_ in (do, dd):
                if d_['dets'].shape[0]:
                    cands.append((abs(float(d_['scores'][row, lab + 1]) - float(d_['dets'][:, 4].min())),
                                  abs(float(so[row, lab + 1] - sd[row, lab + 1])) * 2, 1e-3, 'top-%d cut by score' % max_num))
    return _verdict(j, cands)


def set_overlap(a, b):
    a, b = set(a), set(b)
    return len(a & b) / max(1, len(a | b))


# ------------------------------------------------------------------------------------------
# one workload, one seed: the full comparison
# ------------------------------------------------------------------------------------------
def window_parity(model, sd, frames, metas, key_dim, head='hrnmp', dev='cuda:0'):
    """Runs the oracle and the device end to end on `frames` ([T,3,H,W] CPU) and returns a dict of index
    agreement figures (see module docstring).  Device calls go through the registered modules' public methods."""
    from hvrnet_b200 import ops
    T = frames.shape[0]
    img_shape = metas[0]['img_shape'][:2]
    out = dict(T=T)
    with torch.no_grad():
        c4_ref = R.trunk_forward(sd, frames)
        cls_o, reg_o = R.rpn_forward(sd, c4_ref)
    # ---- device: trunk per frame through the reference call surface, RPN maps, proposals with anchor indices
    c4s = [model(img=frames[i:i + 1].to(dev), img_meta=[metas[i]], backbone_feat=True)[0] for i in range(T)]
    c4 = model._window_split(c4s)
    maps = model.rpn_head.forward_maps(c4)
    props_d, counts_d, idx_d = model.rpn_head.proposals_from_maps(maps, metas[0]['img_shape'], model.test_cfg.rpn,
                                                                  want_idx=True)
    counts = counts_d.cpu().tolist()
    cls_d, reg_d = device_maps_to_oracle_layout(maps)
    out['rpn_logit_err'] = float((cls_d - cls_o).abs().max())
    out['rpn_logit_rel'] = float((cls_d - cls_o).abs().max() / cls_o.abs().max())
    anchors = anchors_for(cls_o.shape[-2], cls_o.shape[-1])
    exact_replay = exact_free = 0
    overlap, diverg, props_o = [], [], []
    top_exact = 0
    for t in range(T):
        to = proposal_trace(cls_o[t], reg_o[t], anchors, img_shape)
        td = proposal_trace(cls_d[t], reg_d[t], anchors, img_shape)
        got = idx_d[t, :counts[t]].cpu().tolist()
        # (A) replay: oracle logic on the device's maps == the device's own anchor indices, in order
        exact_replay += int(got == td['anchor'].tolist())
        # device boxes against the replay's boxes (same indices): decode arithmetic
        if got == td['anchor'].tolist() and counts[t]:
            out['replay_box_err_px'] = max(out.get('replay_box_err_px', 0.0),
                                           float((props_d[t, :counts[t], :4].cpu() - td['props'][:, :4]).abs().max()))
        # (B) free-running: oracle on its own maps
        exact_free += int(got == to['anchor'].tolist())
        top_exact += int(to['top'].tolist() == td['top'].tolist())
        overlap.append(set_overlap(got, to['anchor'].tolist()))
        e = explain_proposal_divergence(to, td)
        if e is not None:
            e['frame'] = t
            diverg.append(e)
        props_o.append(to['props'])
    out.update(frames_replay_exact=exact_replay, frames_free_exact=exact_free, frames_top_exact=top_exact,
               proposal_set_overlap_min=min(overlap), proposal_set_overlap_mean=float(np.mean(overlap)),
               proposal_divergences=diverg)
    # ---- second stage on identical rois (the oracle's proposals forced in on both sides)
    frcnn = head == 'shared_fc'
    with torch.no_grad():
        if frcnn:
            c5_ref = R.c5_forward(sd, c4_ref)
            rois_o = R.bbox2roi(props_o)
            feats = cref.roi_align(c5_ref, rois_o)
            c_, r_ = R.shared_fc_forward(sd, feats)
            raux = dict(cls=[c_], reg=[r_], proposals=props_o, length=rois_o.shape[0])
        else:
            _, raux = R.hnmb_forward_feat(sd, list(c4_ref.split(1)), metas, key_dim, head=head,
                                          roi_align_fn=cref.roi_align, return_aux=True)

    def device_stage2(proposals):
        """-> (key rois [n,5] with batch index 0, [cls...], [reg...]) through the detector's own methods"""
        if frcnn:
            _, aux = model._detect(c4, metas, 1, 1, 0, False, proposals=proposals, return_aux=True)
        else:
            _, aux = model(x=c4s, img=None, img_meta=metas, forward_feat=True, return_loss=False, rescale=True,
                           proposals=proposals, return_aux=True)
        s_, n_ = aux['start'], aux['length']
        rk = aux['rois'][s_:s_ + n_].clone()
        rk[:, 0] = 0
        return rk, aux['cls'], aux['reg']

    _, cls_f, reg_f = device_stage2([p.to(dev) for p in raux['proposals']])
    rois_key = torch.cat([torch.zeros(raux['length'], 1), raux['proposals'][key_dim][:, :4]], 1)
    m0 = metas[0]
    rescale = not frcnn                     # the window detectors are driven with rescale=True (tools/hnl_test.py)
    det_exact_replay = det_exact_forced = 0
    det_div = []
    n_out = len(raux['cls'])

    def device_dets(rk, c, r):
        d, l, k, ridx = ops.det_postprocess_batched(rk, c, r, 1, img_shape, m0['scale_factor'], rescale,
                                                    model.bbox_head.target_stds, n_cls=model.bbox_head.num_classes,
                                                    want_idx=True)
        kk = int(k.item())
        return l[0, :kk].cpu().tolist(), ridx[0, :kk].cpu().tolist()

    for o in range(n_out):
        cd, rd = cls_f[o].float().cpu(), reg_f[o].float().cpu()
        out['head_rel_%d' % o] = max(float((cd - raux['cls'][o]).abs().max() / raux['cls'][o].abs().max()),
                                     float((rd - raux['reg'][o]).abs().max() / raux['reg'][o].abs().max()))
        do = det_trace(rois_key, raux['cls'][o], raux['reg'][o], img_shape, m0['scale_factor'], rescale)
        dd = det_trace(rois_key, cd, rd, img_shape, m0['scale_factor'], rescale)
        got = list(zip(*device_dets(rois_key.to(dev), cls_f[o], reg_f[o])))
        det_exact_replay += int(got == list(zip(dd['labels'].tolist(), dd['rows'].tolist())))
        det_exact_forced += int(got == list(zip(do['labels'].tolist(), do['rows'].tolist())))
        e = explain_det_divergence(do, dd)
        if e is not None:
            e['output'] = o
            det_div.append(e)
        out['det_set_overlap_forced_%d' % o] = set_overlap(got, list(zip(do['labels'].tolist(), do['rows'].tolist())))
    out.update(n_outputs=n_out, det_replay_exact=det_exact_replay, det_forced_exact=det_exact_forced,
               det_divergences=det_div)
    # ---- free-running end to end: detections identified by (label, anchor index of their proposal)
    rk, cls_e, reg_e = device_stage2(None)
    key_anchor_d = idx_d[key_dim, :counts[key_dim]].cpu().tolist()
    key_anchor_o = proposal_trace(cls_o[key_dim], reg_o[key_dim], anchors, img_shape)['anchor'].tolist()
    for o in range(n_out):
        lab, ridx = device_dets(rk, cls_e[o], reg_e[o])
        got = [(a, key_anchor_d[b]) for a, b in zip(lab, ridx)]
        do = det_trace(rois_key, raux['cls'][o], raux['reg'][o], img_shape, m0['scale_factor'], rescale)
        want = [(a, key_anchor_o[b]) for a, b in zip(do['labels'].tolist(), do['rows'].tolist())]
        out['det_free_exact_%d' % o] = int(got == want)
        out['det_free_set_overlap_%d' % o] = set_overlap(got, want)
        out['det_free_count_%d' % o] = (len(got), len(want))
    return out


def format_report(name, r):
    lines = ['%s: T=%d' % (name, r['T'])]
    lines.append('  RPN logits: max |device - oracle| = %.3e (%.2e of the range)' % (r['rpn_logit_err'], r['rpn_logit_rel']))
    lines.append('  (A) replay on the device maps : %d / %d frames with identical anchor-index lists (ordered); '
                 'proposal boxes within %.2e px' % (r['frames_replay_exact'], r['T'], r.get('replay_box_err_px', 0.0)))
    lines.append('  (B) free-running vs the oracle: %d / %d frames identical; pre-NMS top-6000 order identical in %d / %d; '
                 'anchor-set overlap min %.4f mean %.4f' % (r['frames_free_exact'], r['T'], r['frames_top_exact'], r['T'],
                                                             r['proposal_set_overlap_min'], r['proposal_set_overlap_mean']))
    for e in r['proposal_divergences']:
        lines.append('      frame %2d first divergence at kept position %3d: %s; margin %.3e vs arithmetic error %.3e -> %s'
                     % (e['frame'], e['position'], e['what'], e['margin'], e['error'],
                        'near-tie' if e['near_tie'] else 'NOT EXPLAINED'))
    lines.append('  second stage on the oracle\'s proposals: head outputs within %s of the oracle'
                 % ', '.join('%.2e' % r['head_rel_%d' % o] for o in range(r['n_outputs'])))
    lines.append('  (A) replay on the device head outputs: %d / %d outputs with identical (label, roi) lists (ordered)'
                 % (r['det_replay_exact'], r['n_outputs']))
    lines.append('  (B) vs the oracle on its own head outputs: %d / %d identical; set overlap %s'
                 % (r['det_forced_exact'], r['n_outputs'],
                    ', '.join('%.4f' % r['det_set_overlap_forced_%d' % o] for o in range(r['n_outputs']))))
    for e in r['det_divergences']:
        lines.append('      output %d first divergence at position %3d: %s; margin %.3e vs arithmetic error %.3e -> %s'
                     % (e['output'], e['position'], e['what'], e['margin'], e['error'],
                        'near-tie' if e['near_tie'] else 'NOT EXPLAINED'))
    lines.append('  free-running end to end, detections as (label, anchor of the proposal): identical %s; set overlap %s'
                 % (', '.join(str(r['det_free_exact_%d' % o]) for o in range(r['n_outputs'])),
                    ', '.join('%.4f' % r['det_free_set_overlap_%d' % o] for o in range(r['n_outputs']))))
    return '\n'.join(lines)
