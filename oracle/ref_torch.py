"""CPU oracle for the HVRNet per-key-frame inference hot path (TEST INFRASTRUCTURE ONLY).

This file restates, in plain functional torch fp32 on the CPU, the arithmetic the
reference (youthHan/HVRNet, an mmdetection-v1 fork) performs on the path
    R101-C4 trunk -> C5 shared head -> RPN / proposals -> RoIAlign ->
    hierarchical SELSA-style relation head -> box decode -> multiclass NMS.
It is the checker for the CUDA product in ``hvrnet_b200``; nothing in the product
imports it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
cpu_baseline / ``--impl reference`` leg may import this package.

Parity status (tests/test_oracle_golden.py).  Pinned against the reference's OWN code:
anchors, delta2bbox, bbox2roi, multiclass_nms and eval_map (its pure-Python files loaded
standalone, tests/golden/make_golden.py); NMS (its nms_cpu.cpp compiled unmodified,
oracle/_ref, and its doctests); both relation heads - HRNMPBBoxHead.forward_test and
SelsaBBoxHead.forward with forward_single_selsa and _add_selsa_with_fc - and the proposal
generation RPNHead.get_bboxes_single (the methods cut out of the files' syntax trees and
run on bare nn.Module harnesses, tests/golden/make_heads_golden.py: heads to 1e-6,
proposals bit for bit); the video descriptor / similarity of get_triplet_patches
(tests/golden/make_triplet_golden.py).  The reference package itself cannot be imported
(missing mmcv / private pytorch_metric_learning fork / three missing head modules / a broken
unpacking in HRNMPBBoxHead.__init__).
The convolution stacks (the reference's real ResNet trunk, ResLayer shared head and
RPNHead.forward_single, tests/golden/make_convs_golden.py: identical bits), the decode +
multiclass NMS step (HRNMPBBoxHead.get_det_bboxes) and the detector's control flow end to
end (HNMBRCNN.forward_feat / simple_test_bboxes with the oracle's C RoIAlign as the only
substituted piece, tests/golden/make_e2e_golden.py) are pinned the same way.
RoIAlign forward values and the GPU NMS semantic are pinned on the GPU box by the reference's
own CUDA ops compiled unmodified into oracle/_ref (tests/test_gpu_kernels.py::
test_roi_align_vs_reference_cuda_op, ::test_nms_vs_reference_cuda_op).
Still **parity unpinned** by the reference: the inference-time inter-video stage
(oracle-defined; the reference has it in training only).

Every function is keyed on a flat ``state_dict`` that uses the reference's parameter
names, so the same weights load into the oracle and into the CUDA modules.

Repairs made while restating, written down once here:
 (1) the HRNMP head has 4 relation stages (what ``_add_selsa_with_fc`` builds,
     hrnmp_bbox_head.py:134-189) - the 6-way unpack at :100-103 is ignored;
 (2) the SELSA test path takes (cls, reg, _) from the 3-tuple (selsa_rcnn.py:306
     vs selsa_bbox_head.py:261);
 (3) ``cur_range.start`` is an int (hnmb_rcnn.py:586 uses np.sum([]) -> float);
 (4) RoIAlign has no CPU implementation in the reference (roi_align.py:24-28); it
     is restated from the CUDA kernel roi_align_kernel.cu:16-118 in strict IEEE
     fp32 without fused multiply-add contraction;
 (5) NMS threshold semantic: canonical is the GPU kernel's strict ``>``
     (nms_kernel.cu:61); the CPU file's ``>=`` (nms_cpu.cpp:55) is available as
     ``strict_gt=False`` for cross-checks;
 (6) every sort / top-k uses a total order: key descending, then original index
     ascending (the reference's sorts are unstable, rpn_head.py:78,
     nms_kernel.cu:79, bbox_nms.py:58).  The RPN top-k orders by the raw logit,
     which refines the order by sigmoid score (sigmoid is monotone) and is one of
     the outcomes the reference's unstable top-k may produce.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# total-order helpers
# ----------------------------------------------------------------------------


def argsort_desc_stable(keys):
    """Indices ordering ``keys`` descending, ties by ascending index (repair 6)."""
    return torch.sort(keys, descending=True, stable=True)[1]


# ----------------------------------------------------------------------------
# R4  anchors            mmdet/core/anchor/anchor_generator.py:29-83
# ----------------------------------------------------------------------------


def gen_base_anchors(base_size, scales, ratios):
    """Ratio-major, scale-minor base anchors, rounded (anchor_generator.py:29-56)."""
    scales = torch.tensor(scales, dtype=torch.float32)
    ratios = torch.tensor(ratios, dtype=torch.float32)
    w = h = float(base_size)
    xc = 0.5 * (w - 1)
    yc = 0.5 * (h - 1)
    hr = torch.sqrt(ratios)
    wr = 1 / hr
    ws = (w * wr[:, None] * scales[None, :]).reshape(-1)
    hs = (h * hr[:, None] * scales[None, :]).reshape(-1)
    return torch.stack([xc - 0.5 * (ws - 1), yc - 0.5 * (hs - 1),
                        xc + 0.5 * (ws - 1), yc + 0.5 * (hs - 1)], dim=-1).round()


def grid_anchors(base_anchors, featmap_size, stride):
    """All anchors of a (h, w) map, order (y, x, a) (anchor_generator.py:66-83)."""
    fh, fw = featmap_size
    sx = torch.arange(0, fw) * stride
    sy = torch.arange(0, fh) * stride
    xx = sx.repeat(fh)
    yy = sy.view(-1, 1).repeat(1, fw).view(-1)
    shifts = torch.stack([xx, yy, xx, yy], dim=-1).to(base_anchors.dtype)
    return (base_anchors[None] + shifts[:, None]).reshape(-1, 4)


# ----------------------------------------------------------------------------
# R6  delta2bbox         mmdet/core/bbox/transforms.py:34-111
# ----------------------------------------------------------------------------

WH_RATIO_CLIP = 16 / 1000


def delta2bbox(rois, deltas, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.),
               max_shape=None, wh_ratio_clip=WH_RATIO_CLIP):
    """Class-agnostic (N,4) decode.  px + pw*dx is evaluated as addcmul (one
    rounding on CPU builds with FMA, like the reference's torch.addcmul at :98-99)."""
    means = deltas.new_tensor(means)
    stds = deltas.new_tensor(stds)
    d = deltas * stds + means
    dx, dy, dw, dh = d[:, 0], d[:, 1], d[:, 2], d[:, 3]
    max_ratio = abs(math.log(wh_ratio_clip))
    dw = dw.clamp(-max_ratio, max_ratio)
    dh = dh.clamp(-max_ratio, max_ratio)
    px = (rois[:, 0] + rois[:, 2]) * 0.5
    py = (rois[:, 1] + rois[:, 3]) * 0.5
    pw = rois[:, 2] - rois[:, 0] + 1.0
    ph = rois[:, 3] - rois[:, 1] + 1.0
    gw = pw * dw.exp()
    gh = ph * dh.exp()
    gx = torch.addcmul(px, pw, dx)
    gy = torch.addcmul(py, ph, dy)
    x1 = gx - gw * 0.5 + 0.5
    y1 = gy - gh * 0.5 + 0.5
    x2 = gx + gw * 0.5 - 0.5
    y2 = gy + gh * 0.5 - 0.5
    if max_shape is not None:
        x1 = x1.clamp(0, max_shape[1] - 1)
        y1 = y1.clamp(0, max_shape[0] - 1)
        x2 = x2.clamp(0, max_shape[1] - 1)
        y2 = y2.clamp(0, max_shape[0] - 1)
    return torch.stack([x1, y1, x2, y2], dim=-1)


# ----------------------------------------------------------------------------
# R7  NMS                mmdet/ops/nms/src/nms_kernel.cu:14-22,57-63,116-135
#                        mmdet/ops/nms/src/nms_cpu.cpp:5-59
# ----------------------------------------------------------------------------


def _iou_row(box, boxes):
    """IoU of one box against many, +1 pixel convention, fp32, same operation order
    as devIoU (nms_kernel.cu:14-22): interS / (Sa + Sb - interS)."""
    left = torch.maximum(box[0], boxes[:, 0])
    right = torch.minimum(box[2], boxes[:, 2])
    top = torch.maximum(box[1], boxes[:, 1])
    bottom = torch.minimum(box[3], boxes[:, 3])
    w = (right - left + 1).clamp(min=0)
    h = (bottom - top + 1).clamp(min=0)
    inter = w * h
    sa = (box[2] - box[0] + 1) * (box[3] - box[1] + 1)
    sb = (boxes[:, 2] - boxes[:, 0] + 1) * (boxes[:, 3] - boxes[:, 1] + 1)
    return inter / (sa + sb - inter)


def nms(dets, iou_thr, strict_gt=True, max_keep=None):
    """Greedy NMS.  dets (n,5) fp32 [x1,y1,x2,y2,score].  Returns kept indices into
    ``dets`` in ASCENDING original-index order (nms_kernel.cu:132-135,
    nms_cpu.cpp:58).  ``max_keep`` stops after that many survivors *in score order*
    (used only to model ``proposals[:nms_post]`` cheaply; None = reference)."""
    n = dets.shape[0]
    if n == 0:
        return torch.zeros(0, dtype=torch.long)
    order = argsort_desc_stable(dets[:, 4])
    b = dets[order, :4].contiguous()
    removed = torch.zeros(n, dtype=torch.bool)
    keep = []
    for i in range(n):
        if removed[i]:
            continue
        keep.append(i)
        if max_keep is not None and len(keep) >= max_keep:
            break
        if i + 1 < n:
            iou = _iou_row(b[i], b[i + 1:])
            removed[i + 1:] |= (iou > iou_thr) if strict_gt else (iou >= iou_thr)
    kept = order[torch.tensor(keep, dtype=torch.long)]
    if max_keep is None:
        kept = torch.sort(kept)[0]
    return kept


# ----------------------------------------------------------------------------
# R8  RoIAlign           mmdet/ops/roi_align/src/roi_align_kernel.cu:16-118
# ----------------------------------------------------------------------------


def roi_align(feat, rois, out_size=7, spatial_scale=1 / 16., sample_num=2):
    """feat (B,C,H,W) fp32, rois (n,5) [batch,x1,y1,x2,y2] -> (n,C,ph,pw).

    Vectorised over channels; the per-sample arithmetic and the accumulation order
    (iy outer, ix inner; w1*lt + w2*rt + w3*lb + w4*rb left to right; divide by the
    sample count last) follow roi_align_kernel.cu:16-61,102-117 with every product
    and sum rounded separately (no FMA) - torch elementwise ops round each step."""
    B, C, H, W = feat.shape
    ph = pw = int(out_size)
    n = rois.shape[0]
    out = feat.new_zeros((n, C, ph, pw))
    f32 = torch.float32
    scale = torch.tensor(spatial_scale, dtype=f32)
    for r in range(n):
        b = int(rois[r, 0])
        x1 = rois[r, 1] * scale
        y1 = rois[r, 2] * scale
        x2 = (rois[r, 3] + 1) * scale
        y2 = (rois[r, 4] + 1) * scale
        rw = torch.clamp(x2 - x1, min=0.)
        rh = torch.clamp(y2 - y1, min=0.)
        bh = rh / ph
        bw = rw / pw
        sh = sample_num if sample_num > 0 else int(math.ceil(float(rh) / ph))
        sw = sample_num if sample_num > 0 else int(math.ceil(float(rw) / pw))
        fm = feat[b]
        for p in range(ph):
            for q in range(pw):
                acc = torch.zeros(C, dtype=f32)
                for iy in range(sh):
                    y = y1 + p * bh + torch.tensor(iy + .5, dtype=f32) * bh / torch.tensor(float(sh), dtype=f32)
                    for ix in range(sw):
                        x = x1 + q * bw + torch.tensor(ix + .5, dtype=f32) * bw / torch.tensor(float(sw), dtype=f32)
                        acc = acc + _bilinear(fm, H, W, y.clone(), x.clone())
                out[r, :, p, q] = acc / float(sh * sw)
    return out


def _bilinear(fm, H, W, y, x):
    """bilinear_interpolate (roi_align_kernel.cu:16-61) for all channels."""
    C = fm.shape[0]
    if y < -1.0 or y > H or x < -1.0 or x > W:
        return torch.zeros(C, dtype=torch.float32)
    if y <= 0:
        y = torch.zeros((), dtype=torch.float32)
    if x <= 0:
        x = torch.zeros((), dtype=torch.float32)
    yl = int(y)
    xl = int(x)
    if yl >= H - 1:
        yh = yl = H - 1
        y = torch.tensor(float(yl), dtype=torch.float32)
    else:
        yh = yl + 1
    if xl >= W - 1:
        xh = xl = W - 1
        x = torch.tensor(float(xl), dtype=torch.float32)
    else:
        xh = xl + 1
    ly = y - yl
    lx = x - xl
    hy = 1. - ly
    hx = 1. - lx
    w1, w2, w3, w4 = hy * hx, hy * lx, ly * hx, ly * lx
    return w1 * fm[:, yl, xl] + w2 * fm[:, yl, xh] + w3 * fm[:, yh, xl] + w4 * fm[:, yh, xh]


# ----------------------------------------------------------------------------
# R1  trunk              mmdet/models/backbones/resnet.py:222-257,269-329,522-533
# R2  C5 shared head     mmdet/models/shared_heads/res_layer.py:16-52,67-74
# ----------------------------------------------------------------------------

BN_EPS = 1e-5  # mmdet/models/utils/norm.py:44


def _bn(sd, name, x):
    return F.batch_norm(x, sd[name + '.running_mean'], sd[name + '.running_var'],
                        sd[name + '.weight'], sd[name + '.bias'], False, 0., BN_EPS)


def bottleneck(sd, p, x, stride, dilation, has_down):
    """Caffe-style bottleneck: the stride sits on the first 1x1 (resnet.py:129-134);
    3x3 uses padding=dilation (:153-161); downsample = 1x1 stride conv + BN
    (:283-294); residual add then ReLU (:251-264)."""
    out = F.relu(_bn(sd, p + 'bn1', F.conv2d(x, sd[p + 'conv1.weight'], stride=stride)))
    out = F.relu(_bn(sd, p + 'bn2', F.conv2d(out, sd[p + 'conv2.weight'], padding=dilation,
                                             dilation=dilation)))
    out = _bn(sd, p + 'bn3', F.conv2d(out, sd[p + 'conv3.weight']))
    idt = x
    if has_down:
        idt = _bn(sd, p + 'downsample.1', F.conv2d(x, sd[p + 'downsample.0.weight'], stride=stride))
    return F.relu(out + idt)


R101_BLOCKS = (3, 4, 23, 3)  # resnet.py:377


def res_layer(sd, p, x, blocks, stride, dilation):
    for i in range(blocks):
        x = bottleneck(sd, '%s%d.' % (p, i), x, stride if i == 0 else 1, dilation, i == 0)
    return x


def trunk_forward(sd, img, prefix='backbone.', strides=(1, 2, 2), dilations=(1, 1, 1)):
    """ResNet-101 conv1..layer3 -> C4 (resnet.py:522-533 with hrnmp cfg:39-50)."""
    x = F.conv2d(img, sd[prefix + 'conv1.weight'], stride=2, padding=3)
    x = F.relu(_bn(sd, prefix + 'bn1', x))
    x = F.max_pool2d(x, 3, 2, 1)
    for i in range(len(strides)):
        x = res_layer(sd, '%slayer%d.' % (prefix, i + 1), x, R101_BLOCKS[i], strides[i], dilations[i])
    return x


def c5_forward(sd, c4, prefix='shared_head.', stride=1, dilation=2):
    """layer4 on the whole C4 map then new_layer_1 = 1x1 conv(+bias) + ReLU, no norm
    (res_layer.py:67-74, conv_module.py:95-97,156-164)."""
    x = res_layer(sd, prefix + 'layer4.', c4, R101_BLOCKS[3], stride, dilation)
    x = F.conv2d(x, sd[prefix + 'new_layer_1.conv.weight'], sd[prefix + 'new_layer_1.conv.bias'])
    return F.relu(x)


# ----------------------------------------------------------------------------
# R3  RPN head           mmdet/models/anchor_heads/rpn_head.py:30-35
# R5  proposals          rpn_head.py:55-104, anchor_head.py:209-278
# ----------------------------------------------------------------------------


def rpn_forward(sd, c4, prefix='rpn_head.'):
    x = F.relu(F.conv2d(c4, sd[prefix + 'rpn_conv.weight'], sd[prefix + 'rpn_conv.bias'], padding=1))
    cls = F.conv2d(x, sd[prefix + 'rpn_cls.weight'], sd[prefix + 'rpn_cls.bias'])
    reg = F.conv2d(x, sd[prefix + 'rpn_reg.weight'], sd[prefix + 'rpn_reg.bias'])
    return cls, reg


def rpn_proposals_single(cls, reg, anchors, img_shape, nms_pre=6000, nms_post=300, max_num=300,
                         nms_thr=0.7, return_aux=False, min_bbox_size=0, strict_gt=True):
    """One frame.  cls (A,h,w) sigmoid logits, reg (4A,h,w).  Returns (k,5).

    permute(1,2,0) flatten -> sigmoid -> top nms_pre (sorted) -> decode (stds 1,
    clamp to img_shape) -> [min_bbox_size filter, :84-90; 0 in the configs] -> NMS -> first
    nms_post -> top max_num by score (rpn_head.py:63-103).  Pinned against the reference's own
    get_bboxes_single by tests/golden/ref_heads_golden.pt (strict_gt=False: its CPU NMS)."""
    logits = cls.permute(1, 2, 0).reshape(-1)
    deltas = reg.permute(1, 2, 0).reshape(-1, 4)
    if nms_pre > 0 and logits.shape[0] > nms_pre:
        top = argsort_desc_stable(logits)[:nms_pre]
    else:
        top = torch.arange(logits.shape[0])      # the reference leaves the order alone here
    scores = logits[top].sigmoid()
    boxes = delta2bbox(anchors[top], deltas[top], max_shape=img_shape)
    if min_bbox_size > 0:
        w, h = boxes[:, 2] - boxes[:, 0] + 1, boxes[:, 3] - boxes[:, 1] + 1
        valid = torch.nonzero((w >= min_bbox_size) & (h >= min_bbox_size)).reshape(-1)
        top, scores, boxes = top[valid], scores[valid], boxes[valid]
    dets = torch.cat([boxes, scores[:, None]], dim=-1)
    # NMS sorts by its 5th column; hand it the logit so that saturated-sigmoid ties keep
    # the (logit desc, index asc) total order of repair 6.
    keep = nms(torch.cat([boxes, logits[top][:, None]], dim=-1), nms_thr, strict_gt=strict_gt)
    keep = keep[argsort_desc_stable(logits[top][keep])]   # identity when `top` is sorted
    keep = keep[:nms_post]
    keep = keep[:min(max_num, keep.shape[0])]             # final top-k (:100-103): same order
    props = dets[keep]
    if return_aux:
        return props, top[keep]
    return props


def rpn_proposals(sd_or_outs, c4=None, img_shapes=None, cfg=None, prefix='rpn_head.',
                  anchor_scales=(4, 8, 16, 32), anchor_ratios=(0.5, 1.0, 2.0), stride=16):
    """All frames of a window (anchor_head.py:254-278)."""
    cls, reg = rpn_forward(sd_or_outs, c4, prefix) if c4 is not None else sd_or_outs
    cfg = dict(nms_pre=6000, nms_post=300, max_num=300, nms_thr=0.7) if cfg is None else cfg
    base = gen_base_anchors(stride, anchor_scales, anchor_ratios)
    anchors = grid_anchors(base, cls.shape[-2:], stride)
    return [rpn_proposals_single(cls[i], reg[i], anchors, img_shapes[i][:2], **cfg)
            for i in range(cls.shape[0])]


# ----------------------------------------------------------------------------
# R9  relation operator  mmdet/models/bbox_heads/hrnmp_bbox_head.py:216-355
# ----------------------------------------------------------------------------


def relation(sd, p, idx, X, q_range=None):
    """SELSA-style non-local block ``idx`` on X (N,D).  Keys/values are all rows
    (:249-250); queries are all rows, or the rows of ``q_range=(start,len)`` when
    idx_output_cur_only (:269-279).  Q,K are linear projections (:283-288), logits
    are scaled by 1/sqrt(dim[1]) (:293-294), softmax over keys (:332), V is the
    un-projected input (conv_g False, :289,:340-342), output through the 1x1
    Conv2d ``linear_out`` (:343-350)."""
    s = '%sselsa_%d.' % (p, idx)
    Xq = X if q_range is None else X[q_range[0]:q_range[0] + q_range[1]]
    Q = F.linear(Xq, sd[s + 'q_data_fc_%d.weight' % idx], sd[s + 'q_data_fc_%d.bias' % idx])
    K = F.linear(X, sd[s + 'k_data_fc_%d.weight' % idx], sd[s + 'k_data_fc_%d.bias' % idx])
    aff = torch.mm(Q, K.t()) * (1.0 / math.sqrt(float(K.shape[1])))
    P = torch.softmax(aff, dim=1)
    O = torch.mm(P, X)
    Wz = sd[s + 'linear_out_%d.weight' % idx]
    return F.linear(O, Wz.reshape(Wz.shape[0], -1), sd[s + 'linear_out_%d.bias' % idx])


def _fc(sd, name, x):
    return F.linear(x, sd[name + '.weight'], sd[name + '.bias'])


# ----------------------------------------------------------------------------
# R10 HRNMP forward_test hrnmp_bbox_head.py:800-909
# ----------------------------------------------------------------------------


def hrnmp_forward_test(sd, roi_feats, start, length, prefix='bbox_head.', support_rows=None,
                       return_feats=False):
    """roi_feats (N,256,7,7); key rows = [start, start+length).

    Returns ([cls_branch, cls], [reg_branch, reg]).  ``support_rows`` (M,1024), if
    given, are extra *post-fc_new_4* rows appended to the key/value set of stage 4
    (the inter-video definition of SURVEY.md section 8d config 4 - oracle-defined,
    parity unpinned by the reference; None reproduces forward_test exactly)."""
    p = prefix
    s, e = start, start + length
    x = roi_feats.reshape(roi_feats.shape[0], -1)
    f1 = _fc(sd, p + 'fc_new_1', x)                                   # :827-828
    a1 = F.relu(f1 + relation(sd, p, 1, f1))                          # :829-834
    f2 = _fc(sd, p + 'fc_new_2', a1)                                  # :837
    a2 = F.relu(f2 + relation(sd, p, 2, f2))                          # :838-854
    a2k = a2[s:e]
    cls_b = _fc(sd, p + 'fc_cls', a2k)                                # :858-861
    reg_b = _fc(sd, p + 'fc_reg', a2k)
    x3 = torch.cat([f1[:s], a2k, f1[e:]], dim=0)                      # :865-868
    f3 = _fc(sd, p + 'fc_new_3', x3)                                  # :869
    a3 = F.relu(f3 + relation(sd, p, 3, f3))                          # :870-883
    f4 = _fc(sd, p + 'fc_new_4', a3)                                  # :889
    if support_rows is None:
        att4 = relation(sd, p, 4, f4, (s, length))                    # :890-891
    else:
        att4 = relation_with_support(sd, p, 4, f4, (s, length), support_rows)
    a4 = F.relu(f4[s:e] + att4)                                       # :892-903
    cls = _fc(sd, p + 'fc_cls_2', a4)                                 # :905-906
    reg = _fc(sd, p + 'fc_reg_2', a4)
    if return_feats:
        return [cls_b, cls], [reg_b, reg], dict(f1=f1, a1=a1, f2=f2, a2=a2, f3=f3, a3=a3, f4=f4, a4=a4)
    return [cls_b, cls], [reg_b, reg]


def relation_with_support(sd, p, idx, X, q_range, support):
    """Stage-4 relation whose key/value set is cat(X, support) (R9x; same math as
    ``relation``; queries are rows of X only)."""
    Xk = torch.cat([X, support], dim=0)
    s = '%sselsa_%d.' % (p, idx)
    Xq = X[q_range[0]:q_range[0] + q_range[1]]
    Q = F.linear(Xq, sd[s + 'q_data_fc_%d.weight' % idx], sd[s + 'q_data_fc_%d.bias' % idx])
    K = F.linear(Xk, sd[s + 'k_data_fc_%d.weight' % idx], sd[s + 'k_data_fc_%d.bias' % idx])
    P = torch.softmax(torch.mm(Q, K.t()) * (1.0 / math.sqrt(float(K.shape[1]))), dim=1)
    O = torch.mm(P, Xk)
    Wz = sd[s + 'linear_out_%d.weight' % idx]
    return F.linear(O, Wz.reshape(Wz.shape[0], -1), sd[s + 'linear_out_%d.bias' % idx])


def hrnmp_stage123_key_feats(sd, roi_feats, start, length, prefix='bbox_head.'):
    """What a support video contributes to another video's stage 4:
    fc_new_4(A3[key]) (hrnmp_bbox_head.py:651-738,742 with mining off)."""
    *_, feats = hrnmp_forward_test(sd, roi_feats, start, length, prefix, return_feats=True)
    return feats['f4'][start:start + length]


# ----------------------------------------------------------------------------
# N4  video-level similarity  hnmb_rcnn.py:76-101 (get_triplet_patches)
# ----------------------------------------------------------------------------


def video_descriptor(c5):
    """c5 (T,C,h,w), the shared head's output for the frames of one video -> (C,):
    global average pool per frame, maximum over the frames (hnmb_rcnn.py:78-81)."""
    return F.adaptive_avg_pool2d(c5, (1, 1)).reshape(c5.shape[0], c5.shape[1]).max(dim=0).values


def video_similarity(queries, candidates):
    """softmax over the candidates of (1/sqrt(C)) * q . c  (hnmb_rcnn.py:85-88, :94-96).
    queries (Q,C), candidates (M,C) -> (Q,M)."""
    scale = 1.0 / math.sqrt(float(queries.shape[-1]))
    return torch.softmax(scale * torch.mm(queries, candidates.t()), dim=1)


def triplet_patches(c5_feats_all, key_video=0, video_per_cls=3):
    """get_triplet_patches (hnmb_rcnn.py:76-101) restated: the first `video_per_cls` videos share the
    key video's class.  Returns [key, the LEAST similar same-class video, the other-class video MOST
    similar to {key, that video} taken together].  Pinned by tests/golden/ref_triplet_golden.pt."""
    d = torch.stack([video_descriptor(c) for c in c5_feats_all])
    same = d[:video_per_cls]
    sim = video_similarity(same[:1], same)                                    # :85-88 (row of video 0)
    hard = int(torch.argmin(sim[:, 1:], dim=1)[0]) + 1                        # :89
    chosen = torch.stack([same[key_video], same[hard]])                       # :91
    other = video_similarity(chosen, d[video_per_cls:]).sum(dim=0, keepdim=True)   # :92-97
    return [key_video, hard, int(torch.argmax(other, dim=1)[0]) + video_per_cls]    # :99-101


def select_support_by_similarity(desc, g, n_support):
    """Inference-time use of the same similarity (oracle-defined, parity unpinned by the reference, which
    selects support videos only in training): the `n_support` videos other than g whose descriptors are
    most similar to video g's - video_similarity of g against all other videos, largest first, ties to
    the lower index.  desc (G,C).  Returns (indices, weights (G-1,) in candidate order)."""
    G = desc.shape[0]
    cand = [i for i in range(G) if i != g]
    if not cand:
        return [], desc.new_zeros(0)
    w = video_similarity(desc[g:g + 1], desc[cand])[0]
    order = sorted(range(len(cand)), key=lambda j: (-float(w[j]), cand[j]))
    return [cand[j] for j in order[:min(n_support, len(cand))]], w


# ----------------------------------------------------------------------------
# R10s SELSA forward     selsa_bbox_head.py:203-261
# ----------------------------------------------------------------------------


def selsa_forward(sd, roi_feats, start, length, prefix='bbox_head.'):
    p = prefix
    x = roi_feats.reshape(roi_feats.shape[0], -1)
    f1 = _fc(sd, p + 'fc_new_1', x)
    a1 = F.relu(f1 + relation(sd, p, 1, f1))
    f2 = _fc(sd, p + 'fc_new_2', a1)
    a2 = f2 + relation(sd, p, 2, f2)
    a2 = F.relu(a2[start:start + length])                             # :250-256
    return _fc(sd, p + 'fc_cls', a2), _fc(sd, p + 'fc_reg', a2)


def shared_fc_forward(sd, roi_feats, prefix='bbox_head.'):
    """SharedFCBBoxHead, 2 fcs (convfc_bbox_head.py:126-167,170-185) - config 1."""
    x = roi_feats.reshape(roi_feats.shape[0], -1)
    x = F.relu(_fc(sd, prefix + 'shared_fcs.0', x))
    x = F.relu(_fc(sd, prefix + 'shared_fcs.1', x))
    return _fc(sd, prefix + 'fc_cls', x), _fc(sd, prefix + 'fc_reg', x)


# ----------------------------------------------------------------------------
# R11 get_det_bboxes     hrnmp_bbox_head.py:1009-1052 / bbox_head.py:132-169
# R12 multiclass_nms     mmdet/core/post_processing/bbox_nms.py:6-66
# R13 bbox2roi/result    mmdet/core/bbox/transforms.py:149-168,181-199
# ----------------------------------------------------------------------------


def multiclass_nms(boxes, scores, score_thr=0.001, iou_thr=0.3, max_num=300, strict_gt=True):
    """boxes (n,4) class-agnostic, scores (n,C) with column 0 = background.
    Returns (k,5), (k,) int64 labels (0-based).  Per class: rows with score > thr
    (:36), NMS with kept rows in ascending row order (nms_wrapper.py:61), label
    c-1; concatenated by class; if more than max_num, the top max_num by score
    (:57-61) under the total order (score desc, concatenated position asc)."""
    dets, labels = [], []
    for c in range(1, scores.shape[1]):
        m = scores[:, c] > score_thr
        if not m.any():
            continue
        d = torch.cat([boxes[m], scores[m, c][:, None]], dim=1)
        keep = nms(d, iou_thr, strict_gt=strict_gt)
        dets.append(d[keep])
        labels.append(torch.full((keep.shape[0],), c - 1, dtype=torch.long))
    if not dets:
        return boxes.new_zeros((0, 5)), torch.zeros(0, dtype=torch.long)
    dets = torch.cat(dets)
    labels = torch.cat(labels)
    if dets.shape[0] > max_num:
        order = argsort_desc_stable(dets[:, 4])[:max_num]
        dets, labels = dets[order], labels[order]
    return dets, labels


def decode_scores_boxes(rois, cls_score, bbox_pred, img_shape, scale_factor=1.0, rescale=False,
                        target_stds=(0.1, 0.1, 0.2, 0.2)):
    scores = F.softmax(cls_score, dim=1)
    boxes = delta2bbox(rois[:, 1:], bbox_pred, (0., 0., 0., 0.), target_stds, img_shape)
    if rescale:
        boxes = boxes / scale_factor
    return boxes, scores


def get_det_bboxes(rois, cls_scores, bbox_preds, img_shape, scale_factor=1.0, rescale=False,
                   score_thr=0.001, iou_thr=0.3, max_per_img=300, strict_gt=True):
    """One (dets, labels) pair per head output (hrnmp_bbox_head.py:1018-1052).  Pinned against the
    reference's own method by tests/golden/ref_heads_golden.pt (strict_gt=False: its CPU NMS)."""
    outs = []
    for c, r in zip(cls_scores, bbox_preds):
        boxes, scores = decode_scores_boxes(rois, c, r, img_shape, scale_factor, rescale)
        outs.append(multiclass_nms(boxes, scores, score_thr, iou_thr, max_per_img, strict_gt=strict_gt))
    return [o[0] for o in outs], [o[1] for o in outs]


def bbox2roi(bbox_list):
    out = []
    for i, b in enumerate(bbox_list):
        if b.shape[0] > 0:
            out.append(torch.cat([b.new_full((b.shape[0], 1), i), b[:, :4]], dim=-1))
        else:
            out.append(b.new_zeros((0, 5)))
    return torch.cat(out, 0)


def bbox2result(dets, labels, num_classes=31):
    if dets.shape[0] == 0:
        return [np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes - 1)]
    d = dets.cpu().numpy()
    lab = labels.cpu().numpy()
    return [d[lab == i, :] for i in range(num_classes - 1)]


# ----------------------------------------------------------------------------
# R14 detector control   detectors/hnmb_rcnn.py:195-222,571-613;
#                        selsa_rcnn.py:56-83,281-317; two_stage.py:280-299
# ----------------------------------------------------------------------------


def hnmb_forward_feat(sd, c4_list, img_metas, key_dim, rescale=True, rpn_cfg=None, rcnn_cfg=None,
                      head='hrnmp', roi_align_fn=None, return_aux=False):
    """One key frame from a window of cached C4 maps.  Returns the list (one entry
    per head output) of per-class result lists, as HNMBRCNN.forward_feat does."""
    roi_align_fn = roi_align if roi_align_fn is None else roi_align_fn
    c4 = torch.cat(tuple(c4_list), dim=0)                              # :200
    c5 = c5_forward(sd, c4)                                            # :202-203
    props = rpn_proposals(sd, c4, [m['img_shape'] for m in img_metas], rpn_cfg)   # :207
    rois_all = [bbox2roi([p]) for p in props]                          # :580-582
    start = int(sum(r.shape[0] for r in rois_all[:key_dim]))           # :586 (repair 3)
    length = rois_all[key_dim].shape[0]
    feats = torch.cat([roi_align_fn(c5[i:i + 1], rois_all[i]) for i in range(len(rois_all))], 0)  # :596-599
    rcnn_cfg = dict(score_thr=0.001, iou_thr=0.3, max_per_img=300) if rcnn_cfg is None else rcnn_cfg
    if head == 'hrnmp':
        cls, reg = hrnmp_forward_test(sd, feats, start, length)        # :602
    else:
        c, r = selsa_forward(sd, feats, start, length)                 # selsa_rcnn.py:306 (repair 2)
        cls, reg = [c], [r]
    m = img_metas[0]                                                   # :603-604 (frame 0's meta)
    dets, labels = get_det_bboxes(rois_all[key_dim], cls, reg, m['img_shape'][:2], m['scale_factor'],
                                  rescale, **rcnn_cfg)
    res = [bbox2result(d, l) for d, l in zip(dets, labels)]
    if return_aux:
        return res, dict(proposals=props, roi_feats=feats, cls=cls, reg=reg, dets=dets, labels=labels,
                         c5=c5, start=start, length=length)
    return res


def faster_rcnn_simple_test(sd, img, img_meta, rescale=False, roi_align_fn=None):
    """Config 1: plain Faster-RCNN R101-C5 with SharedFCBBoxHead on the CPU
    (two_stage.py:280-299, test_mixins.py:40-69)."""
    roi_align_fn = roi_align if roi_align_fn is None else roi_align_fn
    c4 = trunk_forward(sd, img)
    props = rpn_proposals(sd, c4, [img_meta['img_shape']])
    rois = bbox2roi(props)
    c5 = c5_forward(sd, c4)
    feats = roi_align_fn(c5, rois)
    cls, reg = shared_fc_forward(sd, feats)
    dets, labels = get_det_bboxes(rois, [cls], [reg], img_meta['img_shape'][:2], img_meta['scale_factor'], rescale)
    return bbox2result(dets[0], labels[0])
