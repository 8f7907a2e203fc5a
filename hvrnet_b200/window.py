"""The per-key-frame step on fixed-size device buffers: V windows of T frames -> detections of the V key frames.

One code path for the eager detectors (models.py), the CUDA-graph runner (runtime.py) and the inter-video
split: every frame keeps a fixed block of P rows (P = max_num of the RPN test config, or the largest proposal
list handed in by the caller), the per-frame proposal counts stay on the device and act as masks, so neither
launch geometry nor any host decision depends on them - a window whose frames yield FEWER than max_num
proposals (hnmb_rcnn.py:582-599 uses the actual counts) runs the same launches as a full one and a captured
graph replays for it.

Reference control flow mirrored: HNMBRCNN.forward_feat / simple_test_bboxes / get_roi_feat
(hnmb_rcnn.py:195-222, 571-613), SelsaRCNN (selsa_rcnn.py:56-83, 281-317), RPNTestMixin.simple_test_rpn,
BBoxTestMixin (test_mixins.py:9-13, 40-69); the inter-video stage: hrnmp_bbox_head.py:740-795.
"""
import numpy as np
import torch

from . import engine, ops


import os

# experiment switches (bench): HVR_FORK_POST=0 runs the detection post-processing on the main stream instead of a forked
# branch under stages 3-4; HVR_FORK_PROPOSALS=0 does the same for proposal generation (under the C5 convolutions)
FORK_POST = os.environ.get('HVR_FORK_POST', '1') != '0'
FORK_PROPOSALS = os.environ.get('HVR_FORK_PROPOSALS', '1') != '0'


class FrameStages:
    """Outputs of the per-frame stages of V*T frames (C5, RPN, proposals, RoIAlign)."""
    __slots__ = ('maps', 'props', 'counts', 'c5', 'rois', 'rois_key', 'seg', 'key_counts', 'rows', 'P', 'N', 'Npad',
                 'V', 'T')


def props_from_lists(proposals, device):
    """Caller-provided proposals (one [k,5] tensor per frame) -> fixed blocks: props [F,P,5] (zero padded), counts."""
    k = [int(p.shape[0]) for p in proposals]
    P = max(4, ops.round_up(max(k), 4))
    props = torch.zeros((len(proposals), P, 5), dtype=torch.float32, device=device)
    for f, p in enumerate(proposals):
        if k[f]:
            props[f, :k[f], :p.shape[1]] = p.to(device=device, dtype=torch.float32)
    return props, torch.tensor(k, dtype=torch.int32, device=device)


def frame_stages(m, c4, img_shape, V, T, key_dim, perm=None, proposals=None, n_segs=None, side=None):
    """c4 Split NHWC [F,h,w,C] (F = V*T slots) -> FrameStages.  perm: int64 [V*T] buffer slot of window position
    (v, t) (None = identity).  side: optional stream; proposal generation (sort + decode + greedy NMS: latency bound,
    few busy CTAs) then runs on it UNDER the C5 convolutions."""
    fs = FrameStages()
    fs.V, fs.T = V, T
    main = torch.cuda.current_stream()
    if proposals is None:
        fs.maps = m.rpn_head.forward_maps(c4)
        if side is not None:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                fs.props, fs.counts = m.rpn_head.proposals_from_maps(fs.maps, img_shape, m.test_cfg.rpn)
        else:
            fs.props, fs.counts = m.rpn_head.proposals_from_maps(fs.maps, img_shape, m.test_cfg.rpn)
    else:
        fs.maps = None
        fs.props, fs.counts = props_from_lists(proposals, c4.hi.device)
    fs.c5 = m.shared_head.forward_nhwc(c4) if m.feat_from_shared_head else ops.merge(c4)
    if side is not None and proposals is None:
        main.wait_stream(side)
    fs.P = fs.props.shape[1]
    fs.N = T * fs.P
    # relation heads want every video's block 64-row aligned (X^T operands); the plain fc head (Faster-RCNN, T = 1) reads
    # its output rows back as V blocks of P rows, so its blocks are not padded
    pad = m.bbox_head.kind != 'shared_fc'
    fs.Npad = ops.round_up(fs.N, 64) if pad else fs.N
    fs.rois, fs.rois_key, fs.seg, fs.key_counts = ops.window_rois(fs.props, fs.counts, perm, V, T, key_dim, n_segs,
                                                                  pad=pad)
    fs.rows = m.bbox_roi_extractor.roi_layers[0].forward_nhwc_split(fs.c5, fs.rois)
    return fs


class ResultBuffer:
    """One device byte buffer that every result of a step is written into by the producing kernels, so the step ends
    with ONE device->host copy: [per-slot proposal counts int32 F] then per head output: n_dets int32 [V],
    dets f32 [V,M,5], labels int64 [V,M]."""

    def __init__(self, F, V, n_out, M, device):
        self.F, self.V, self.n_out, self.M = F, V, n_out, M
        off = ops.round_up(4 * F, 16)
        self.layout = []
        for _ in range(n_out):
            o_nd = off
            o_d = ops.round_up(o_nd + 4 * V, 16)
            o_l = ops.round_up(o_d + 4 * V * M * 5, 16)
            off = ops.round_up(o_l + 8 * V * M, 16)
            self.layout.append((o_nd, o_d, o_l))
        self.nbytes = off
        self.buf = torch.zeros(off, dtype=torch.uint8, device=device)
        self.counts = self.buf[:4 * F].view(torch.int32)
        self.outs = [(self.buf[o_d:o_d + 4 * V * M * 5].view(torch.float32).view(V, M, 5),
                      self.buf[o_l:o_l + 8 * V * M].view(torch.int64).view(V, M),
                      self.buf[o_nd:o_nd + 4 * V].view(torch.int32)) for o_nd, o_d, o_l in self.layout]

    def parse(self, host):
        """host: uint8 CPU copy of buf -> (counts [F] list, per video: list over outputs of (dets [k,5], labels [k]))."""
        a = host.numpy()
        V, M = self.V, self.M
        counts = a[:4 * self.F].view(np.int32).tolist()
        per_video = [[] for _ in range(V)]
        for o_nd, o_d, o_l in self.layout:
            nd = a[o_nd:o_nd + 4 * V].view(np.int32)
            d = a[o_d:o_d + 4 * V * M * 5].view(np.float32).reshape(V, M, 5)
            lab = a[o_l:o_l + 8 * V * M].view(np.int64).reshape(V, M)
            for v in range(V):
                k = int(nd[v])
                per_video[v].append((torch.from_numpy(d[v, :k].copy()), torch.from_numpy(lab[v, :k].copy())))
        return counts, per_video


def post_process(m, fs, o, V, img_shape, scale_factor, rescale, out=None, want_idx=False):
    """get_det_bboxes (decode + multiclass NMS, hrnmp_bbox_head.py:1009-1052) of one head output o fp32 [V*P, 64] for
    the V key frames: one launch per stage; only the first key_counts[v] rows of a key frame are proposals."""
    head = m.bbox_head
    cls, reg = head._split_out(o)
    cfg = m.test_cfg.rcnn
    nms_cfg = dict(cfg['nms'])
    assert nms_cfg.pop('type', 'nms') == 'nms'
    return ops.det_postprocess_batched(fs.rois_key, cls, reg, V, img_shape[:2], scale_factor, rescale,
                                       head.target_stds, cfg['score_thr'], nms_cfg['iou_thr'], cfg['max_per_img'],
                                       n_cls=head.num_classes, n_valid=fs.key_counts, want_idx=want_idx, out=out)


def head_outputs(m, fs, key_dim):
    """Relation head of the V windows -> list of head outputs fp32 [V*P, 64] ([cls | reg] columns), in the order the
    reference returns them (hrnmp: [branch after stage 2, final after stage 4]; SELSA / shared fc: one)."""
    head = m.bbox_head
    packed = head.packed(fs.rows.hi.device)
    V, P, N, Npad = fs.V, fs.P, fs.N, fs.Npad
    s = key_dim * P
    if head.kind == 'shared_fc':
        assert fs.T == 1 and fs.Npad == P, 'the fc head post-processes V blocks of P rows'
        return [engine.shared_fc_forward(packed, fs.rows)]
    assert head.nongt_dim >= N, 'window rows exceed sampler_num * t_dim (hrnmp_bbox_head.py:249)'
    mask = engine.KeyMask(fs.seg, P)
    if head.kind == 'hrnmp':
        return list(engine.hrnmp_forward_batched(packed, fs.rows, V, N, Npad, s, P, mask=mask))
    return [engine.selsa_forward_batched(packed, fs.rows, V, N, Npad, s, P, mask=mask)]


def scale_of(meta):
    sf = meta['scale_factor']
    return float(sf if not hasattr(sf, '__len__') else np.asarray(sf).reshape(-1)[0])


def detect_windows(m, c4, img_meta, V, T, key_dim, rescale, perm=None, proposals=None, side=None, result=None,
                   keep=None):
    """The whole window stage for V windows.  Returns (ResultBuffer, FrameStages, head outputs).  side: optional list of
    two streams for the forked branches (proposal generation under C5; post-processing of the head outputs)."""
    meta = img_meta[0]                                   # frame 0's meta, hnmb_rcnn.py:603-604
    sf = scale_of(meta)
    main = torch.cuda.current_stream()
    fs = frame_stages(m, c4, meta['img_shape'], V, T, key_dim, perm=perm, proposals=proposals,
                      side=side[0] if (side and FORK_PROPOSALS) else None)
    head = m.bbox_head
    n_out = 2 if head.kind == 'hrnmp' else 1
    if result is None:
        result = ResultBuffer(V * T, V, n_out, m.test_cfg.rcnn['max_per_img'], fs.rows.hi.device)
    forked = []

    def post(j, o):
        st = side[j % 2] if (side and FORK_POST) else None
        if st is None:
            return post_process(m, fs, o, V, meta['img_shape'], sf, rescale, out=result.outs[j])
        st.wait_stream(main)
        forked.append(st)
        with torch.cuda.stream(st):
            return post_process(m, fs, o, V, meta['img_shape'], sf, rescale, out=result.outs[j])

    packed = head.packed(fs.rows.hi.device)
    P, N, Npad = fs.P, fs.N, fs.Npad
    s = key_dim * P
    if head.kind == 'hrnmp':
        # the branch output is post-processed (forked) under stages 3-4
        assert head.nongt_dim >= N, 'window rows exceed sampler_num * t_dim (hrnmp_bbox_head.py:249)'
        mask = engine.KeyMask(fs.seg, P)
        out1, f4, f4T = engine.hrnmp_stage123_batched(packed, fs.rows, V, N, Npad, s, P, mask=mask)
        post(0, out1)
        out2 = engine.hrnmp_stage4_batched(packed, f4, f4T, V, N, Npad, s, P, mask=mask)
        post(1, out2)
        outs = [out1, out2]
        if keep is not None:
            keep += [f4, f4T]
    else:
        outs = head_outputs(m, fs, key_dim)
        post(0, outs[0])
    for st in forked:
        main.wait_stream(st)
    # per-slot proposal counts ride in the same buffer (a 4*F byte device-to-device copy node)
    result.counts.copy_(fs.counts)
    if keep is not None:
        keep += [fs, outs]
    return result, fs, outs


# ------------------------------------------------------------------------------------------------------------
# Inter-video split (BASELINE.json configs 4-5): stage 4 of every local key frame also attends
# to the post-fc_new_4 key rows of n_support other key frames of the whole job.  Three device stages around ONE
# all-gather, each a fixed launch sequence (runtime.GraphRunner captures them as three graphs):
#   A  per-frame stages, relation stages 1-3, fc_new_4, the send buffer of the exchange
#   B  (while the all-gather is in flight) post-processing of the branch output, q_4 / k_4 projections of the own rows
#   C  support rows gathered out of the receive buffer, stage 4 on [own rows | support rows], post-processing
# ------------------------------------------------------------------------------------------------------------
class InterState:
    __slots__ = ('fs', 'out1', 'f4', 'f4k', 'send', 'recv', 'Kown', 'result', 'sel', 'desc', 'rpr', 'n_desc_rows',
                 'S', 'V', 'T', 'key_dim', 'out2', 'sup_rows', 'idx')


def ring_selection(rank, world, V, S, device):
    """int64 [V, S]: ring-order supports (g+1 .. g+S) mod G of the local key frames, -1 where G-1 < S."""
    from .intervideo import support_indices
    G = world * V
    sel = torch.full((V, S), -1, dtype=torch.int64)
    for v in range(V):
        idx = support_indices(rank * V + v, G, S)
        sel[v, :len(idx)] = torch.tensor(idx, dtype=torch.int64)
    return sel.to(device)


def inter_stage_a(m, c4, img_meta, V, T, key_dim, n_support, support_select='ring', perm=None, proposals=None,
                  side=None, result=None):
    from . import intervideo
    meta = img_meta[0]
    st = InterState()
    st.V, st.T, st.key_dim, st.S = V, T, key_dim, n_support
    fs = frame_stages(m, c4, meta['img_shape'], V, T, key_dim, perm=perm, proposals=proposals, n_segs=T + n_support,
                      side=side[0] if side else None)
    st.fs = fs
    head = m.bbox_head
    assert head.kind == 'hrnmp', 'the inter-video stage belongs to the 4-stage HRNMP head'
    dev = fs.rows.hi.device
    if result is None:
        result = ResultBuffer(V * T, V, 2, m.test_cfg.rcnn['max_per_img'], dev)
    st.result = result
    packed = head.packed(dev)
    P, N, Npad = fs.P, fs.N, fs.Npad
    s = key_dim * P
    st.desc = ops.video_descriptor(fs.c5, V) if support_select == 'similarity' else None
    if st.desc is not None and perm is not None:
        raise ValueError('similarity selection needs the frames of a video contiguous in the buffer (no ring)')
    st.out1, st.f4, f4T = engine.hrnmp_stage123_batched(packed, fs.rows, V, N, Npad, s, P,
                                                        mask=engine.KeyMask(fs.seg, P))
    st.f4k = engine.key_rows(st.f4, V, Npad, s, P)
    st.send, st.rpr, st.n_desc_rows = intervideo.pack_exchange(st.f4k, fs.key_counts, st.desc)
    return st


def inter_stage_b(m, st, img_meta, rescale, side=None):
    meta = img_meta[0]
    main = torch.cuda.current_stream()
    fs = st.fs
    packed = m.bbox_head.packed(fs.rows.hi.device)
    if side:
        side[1].wait_stream(main)
        with torch.cuda.stream(side[1]):
            post_process(m, fs, st.out1, st.V, meta['img_shape'], scale_of(meta), rescale, out=st.result.outs[0])
    else:
        post_process(m, fs, st.out1, st.V, meta['img_shape'], scale_of(meta), rescale, out=st.result.outs[0])
    st.Kown, _, _ = engine.lin(st.f4, packed['k4'])
    if side:
        main.wait_stream(side[1])
    return st


def inter_stage_c(m, st, recv, img_meta, rescale, world, rank, sel=None):
    """recv: bf16 [world, 2, rpr, D] (the all-gather's receive buffer; world = 1: the send buffer itself)."""
    from . import intervideo
    meta = img_meta[0]
    fs = st.fs
    V, T, S, P = st.V, st.T, st.S, fs.P
    D = st.f4.shape[1]
    dev = recv.device
    packed = m.bbox_head.packed(dev)
    rpr = st.rpr
    flat = recv.view(world * 2 * rpr, D)
    if st.desc is not None:
        desc_all = intervideo.unpack_descriptors(recv, V, P, st.n_desc_rows, st.desc.shape[1])
        Sq = min(S, world * V - 1)
        sel = torch.full((V, S), -1, dtype=torch.int64, device=dev)
        if Sq > 0:
            sel[:, :Sq] = ops.support_select(desc_all, rank * V, V, Sq)
    elif sel is None:
        sel = ring_selection(rank, world, V, S, dev)
    st.sel = sel
    counts_i32 = flat.view(torch.int32)                   # [rows, D/2]; rank r's counts: row r*2*rpr + V*P
    pool_counts = counts_i32[V * P:].reshape(-1)
    st.idx = ops.support_index(sel, pool_counts, 2 * rpr * (D // 2), V, 2 * rpr, P, T, fs.seg)
    src = ops.Split(flat, flat[rpr:])                     # hi rows / lo rows of a rank are rpr rows apart
    st.sup_rows = ops.Split.empty((V * S * P, D), dev)
    ops.gather_rows(src, st.sup_rows, 1, V * S * P, idx=st.idx, cols=D)
    st.out2 = engine.hrnmp_stage4_inter_batched(packed, st.f4, st.f4k, st.Kown, st.sup_rows, V, fs.N, fs.Npad, P, S * P,
                                                mask=engine.KeyMask(fs.seg, P))
    post_process(m, fs, st.out2, V, meta['img_shape'], scale_of(meta), rescale, out=st.result.outs[1])
    st.result.counts.copy_(fs.counts)
    return st
