"""Thin torch-tensor front-end of the C ABI (include/hvr_b200.h).

torch is used for device memory and streams only; every computation below is a call into
libhvr_b200.so on the current CUDA stream.  CUDA tensors only - there is no CPU path.
"""
import ctypes
import math

import torch

from . import _lib
from ._lib import HvrIGemm, check


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.HvrError('hvrnet_b200 ops take CUDA tensors only (got %s); there is no CPU path' % t.device)


def round_up(v, m):
    return (v + m - 1) // m * m


class Split:
    """Split-bf16 tensor: x ~= hi + lo (two bf16 tensors of identical shape/strides)."""
    __slots__ = ('hi', 'lo')

    def __init__(self, hi, lo):
        self.hi, self.lo = hi, lo

    @staticmethod
    def empty(shape, device):
        return Split(torch.empty(shape, dtype=torch.bfloat16, device=device),
                     torch.empty(shape, dtype=torch.bfloat16, device=device))

    @staticmethod
    def zeros(shape, device):
        return Split(torch.zeros(shape, dtype=torch.bfloat16, device=device),
                     torch.zeros(shape, dtype=torch.bfloat16, device=device))

    @property
    def shape(self):
        return self.hi.shape

    def __getitem__(self, idx):
        return Split(self.hi[idx], self.lo[idx])

    def float(self):
        return merge(self)


def split(x):
    """fp32 tensor (contiguous) -> Split of the same shape."""
    _need_cuda(x)
    x = x.contiguous().float()
    s = Split.empty(x.shape, x.device)
    check(_lib.lib().hvr_split_f32(_p(x), x.numel(), _p(s.hi), _p(s.lo), _stream()), 'hvr_split_f32')
    return s


def split_2d(x, ld_out):
    """fp32 [rows, cols] -> Split [rows, ld_out] (zero padded columns)."""
    _need_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1
    rows, cols = x.shape
    s = Split.empty((rows, ld_out), x.device)
    check(_lib.lib().hvr_split_f32_2d(_p(x), rows, cols, x.stride(0), _p(s.hi), _p(s.lo), ld_out, _stream()),
          'hvr_split_f32_2d')
    return s


def merge(s):
    _need_cuda(s.hi)
    assert s.hi.is_contiguous() and s.lo.is_contiguous()
    out = torch.empty(s.hi.shape, dtype=torch.float32, device=s.hi.device)
    check(_lib.lib().hvr_merge_f32(_p(s.hi), _p(s.lo), s.hi.numel(), _p(out), _stream()), 'hvr_merge_f32')
    return out


def nchw_to_nhwc_split(x):
    _need_cuda(x)
    x = x.contiguous()
    B, C, H, W = x.shape
    s = Split.empty((B, H, W, C), x.device)
    check(_lib.lib().hvr_nchw_to_nhwc_split(_p(x), B, C, H, W, _p(s.hi), _p(s.lo), _stream()), 'hvr_nchw_to_nhwc_split')
    return s


def nhwc_split_to_nchw(s):
    B, H, W, C = s.shape
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=s.hi.device)
    check(_lib.lib().hvr_nhwc_split_to_nchw(_p(s.hi), _p(s.lo), B, C, H, W, _p(out), _stream()),
          'hvr_nhwc_split_to_nchw')
    return out


def nhwc_to_nchw(x):
    B, H, W, C = x.shape
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    check(_lib.lib().hvr_nhwc_to_nchw_f32(_p(x), B, C, H, W, _p(out), _stream()), 'hvr_nhwc_to_nchw_f32')
    return out


def nchw_to_nhwc(x):
    x = x.contiguous()
    B, C, H, W = x.shape
    out = torch.empty((B, H, W, C), dtype=torch.float32, device=x.device)
    check(_lib.lib().hvr_nchw_to_nhwc_f32(_p(x), B, C, H, W, _p(out), _stream()), 'hvr_nchw_to_nhwc_f32')
    return out


# ----------------------------------------------------------------------------------------
# implicit GEMM
# ----------------------------------------------------------------------------------------

def pick_tile(out_w, out_h):
    """M tile (tile_w x tile_h = 128 pixels) wasting the fewest rows."""
    best = None
    for tw in (128, 64, 32, 16, 8):
        th = 128 // tw
        waste = (math.ceil(out_w / tw) * tw) * (math.ceil(out_h / th) * th)
        if best is None or waste < best[0]:
            best = (waste, tw, th)
    return best[1], best[2]


def igemm_desc(a, b, n, *, taps=((0, 0),), a_view=None, out_whb=None, tile=None, alpha=1.0, bias=None, res=None,
               relu=False, out=None, out_f32=None, outT=None, passes=3, b_batch_stride=0, a2=None, a2_view=None):
    """Fill an HvrIGemm.

    a      Split; either 2-D [M, K] (plain GEMM) or NHWC 4-D [B, H, W, C]
    a_view optional (C, W, H, B, stride_w, stride_h, stride_b) override (strided views)
    b      Split [n_rows >= n, ktot] K-major weights / right operand
    out    Split 2-D [rows, ld]; out_f32 fp32 [rows, ld]; outT Split [n, ld]
    """
    g = HvrIGemm()
    g.a_hi, g.a_lo = a.hi.data_ptr(), a.lo.data_ptr()
    if a_view is not None:
        C, W, H, B, sw, sh, sb = a_view
    elif a.hi.dim() == 2:
        M, K = a.shape
        C, W, H, B = K, M, 1, 1
        sw = a.hi.stride(0)
        sh = sw * M
        sb = sh
    else:
        B, H, W, C = a.shape
        sb, sh, sw = a.hi.stride(0), a.hi.stride(1), a.hi.stride(2)
        assert a.hi.stride(3) == 1
    g.a_c, g.a_w, g.a_h, g.a_b = C, W, H, B
    g.a_stride_w, g.a_stride_h, g.a_stride_b = sw, sh, sb
    g.ntaps = len(taps)
    for i, (dx, dy) in enumerate(taps):
        g.tap_dx[i], g.tap_dy[i] = dx, dy
    if out_whb is None:
        out_whb = (W, H, B)
    g.out_w, g.out_h, g.batch = out_whb
    g.tile_w, g.tile_h = tile if tile is not None else pick_tile(g.out_w, g.out_h)
    g.b_hi, g.b_lo, g.n, g.ldb = b.hi.data_ptr(), b.lo.data_ptr(), n, b.hi.stride(0)
    g.alpha = alpha
    g.bias = bias.data_ptr() if bias is not None else None
    if res is not None:
        g.res_hi, g.res_lo, g.ld_res = res.hi.data_ptr(), res.lo.data_ptr(), res.hi.stride(0)
    g.relu = int(relu)
    if out is not None:
        g.out_hi, g.out_lo, g.ld_out = out.hi.data_ptr(), out.lo.data_ptr(), out.hi.stride(0)
    if out_f32 is not None:
        g.out_f32, g.ld_f32 = out_f32.data_ptr(), out_f32.stride(0)
    if outT is not None:
        g.outT_hi, g.outT_lo, g.ld_outT = outT.hi.data_ptr(), outT.lo.data_ptr(), outT.hi.stride(0)
    g.passes = passes
    g.b_stride_batch = b_batch_stride       # elements between the per-image B matrices (0 = one shared B)
    if a2 is not None:                      # second A operand: (C, W, H, B, stride_w, stride_h, stride_b) view
        g.a2_hi, g.a2_lo = a2.hi.data_ptr(), a2.lo.data_ptr()
        g.a2_c, g.a2_w, g.a2_h, g.a2_b, g.a2_stride_w, g.a2_stride_h, g.a2_stride_b = a2_view
    return g


# When set to a list, every hvr_igemm launch is bracketed by CUDA events on the launching
# stream and (start, end, algorithmic_flops) is appended (bench.py's roofline leg).
PROFILE = None


def igemm_flops(g):
    """Algorithmic FLOPs of one descriptor: 2 * rows * n * (ntaps * C)  (FLOP = 2 MAC)."""
    return 2.0 * g.batch * g.out_h * g.out_w * g.n * (g.ntaps * g.a_c + (g.a2_c if g.a2_hi else 0))


def igemm_run(g, check_kernel=False):
    fn = _lib.lib().hvr_igemm_check if check_kernel else _lib.lib().hvr_igemm
    if PROFILE is None:
        check(fn(ctypes.byref(g), _stream()), 'hvr_igemm')
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn(ctypes.byref(g), _stream()), 'hvr_igemm')
    e1.record()
    PROFILE.append((e0, e1, igemm_flops(g), (g.batch * g.out_h * g.out_w, g.n,
                                             g.ntaps * g.a_c + (g.a2_c if g.a2_hi else 0))))


def linear(a, w, n, bias=None, relu=False, res=None, alpha=1.0, want_split=True, want_f32=False, want_T=False,
           passes=3, check_kernel=False, out=None):
    """y = alpha * a @ w[:n].T + bias (+res) (relu).  a Split [M,K]; w Split [>=n, K].
    Returns (Split or None, fp32 or None, Split^T or None).  `out`: caller-provided Split [M, >=n]
    (e.g. a row slice of a batch buffer) instead of a fresh allocation."""
    M = a.shape[0]
    dev = a.hi.device
    ld = round_up(n, 8)
    if out is None:
        out = Split.empty((M, ld), dev) if want_split else None
    of = torch.empty((M, round_up(n, 4)), dtype=torch.float32, device=dev) if want_f32 else None
    oT = None
    if want_T:                                 # only the pad columns need zeros (they meet P's zero padding in P.V;
        ldT = round_up(M, 64)                  # uninitialised memory could hold NaN bit patterns)
        oT = Split.empty((n, ldT), dev)
        if ldT > M:
            oT.hi[:, M:].zero_()
            oT.lo[:, M:].zero_()
    g = igemm_desc(a, w, n, alpha=alpha, bias=bias, res=res, relu=relu, out=out, out_f32=of, outT=oT, passes=passes)
    igemm_run(g, check_kernel)
    return out, of, oT


def transpose_split(s, cols=None):
    """Split [M, ld] (first `cols` columns) -> Split [cols, round_up(M, 64)], pad columns zero."""
    M = s.shape[0]
    cols = cols or s.shape[1]
    out = Split.empty((cols, round_up(M, 64)), s.hi.device)
    check(_lib.lib().hvr_transpose_split(_p(s.hi), _p(s.lo), M, cols, s.hi.stride(0), _p(out.hi), _p(out.lo),
                                         out.hi.stride(0), _stream()), 'hvr_transpose_split')
    return out


def bmm(a, b, V, n, b_batch_stride, alpha=1.0, want_split=True, want_f32=False, out=None, check_kernel=False):
    """torch.bmm over V problems in ONE launch: a Split [V*m, K] (problem v = rows [v*m, (v+1)*m)),
    b Split whose matrix of problem v starts b_batch_stride elements after that of v-1 ([n, K]
    K-major, row stride b.hi.stride(0)).  y_v = alpha * a_v @ b_v[:n].T.  Returns (Split [V*m, ld]
    or None, fp32 [V*m, ld4] or None).  Per problem the same arithmetic as linear()."""
    M, K = a.shape
    assert M % V == 0
    m = M // V
    dev = a.hi.device
    lda = a.hi.stride(0)
    if out is None and want_split:
        out = Split.empty((M, round_up(n, 8)), dev)
    of = torch.empty((M, round_up(n, 4)), dtype=torch.float32, device=dev) if want_f32 else None
    g = igemm_desc(a, b, n, a_view=(K, m, 1, V, lda, lda * m, lda * m), out_whb=(m, 1, V), alpha=alpha,
                   out=out if want_split else None, out_f32=of, b_batch_stride=b_batch_stride)
    igemm_run(g, check_kernel)
    return (out if want_split else None), of


# ----------------------------------------------------------------------------------------
# RoIAlign / NMS / proposals / detections / softmax
# ----------------------------------------------------------------------------------------

def roi_align(feat, rois, out_size=7, spatial_scale=1 / 16., sample_num=2, feat_nhwc=False, out_nhwc=False,
              want_split=False, ld_split=None, want_f32=True, arithmetic='strict'):
    """feat fp32 NCHW (reference layout) or NHWC; rois [n,5].  Returns fp32 output in the
    reference layout [n,C,ph,pw] (or [n,ph,pw,C] when out_nhwc), plus an optional Split
    [n, ld_split] copy in NHWC order.
    arithmetic: 'strict' = every product and sum rounded in the reference's order (bit-exact with the
    reference kernel built with -fmad=false and with the C oracle); 'fast' = the separable FMA evaluation
    (hvr_roi_align_fwd_fast; NHWC in / NHWC out / sample_num 2, else the strict kernels), 1e-5 relative."""
    assert arithmetic in ('strict', 'fast')
    _need_cuda(feat, rois)
    feat = feat.contiguous()
    rois = rois.contiguous().float()
    if feat_nhwc:
        B, H, W, C = feat.shape
    else:
        B, C, H, W = feat.shape
    n = rois.shape[0]
    ph = pw = int(out_size)
    dev = feat.device
    out = None
    if want_f32:
        out = torch.empty((n, ph, pw, C) if out_nhwc else (n, C, ph, pw), dtype=torch.float32, device=dev)
    sp = None
    if want_split:
        assert out_nhwc
        ld_split = ld_split or ph * pw * C
        sp = Split.empty((n, ld_split), dev)
    ws = None if feat_nhwc else torch.empty(feat.numel(), dtype=torch.float32, device=dev)
    if arithmetic == 'fast' and feat_nhwc:
        wsb = _lib.lib().hvr_roi_align_fast_workspace_bytes(n, B)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        check(_lib.lib().hvr_roi_align_fwd_fast(_p(feat), 1, _p(rois), n, B, C, H, W, ph, pw, float(spatial_scale),
                                                int(sample_num), _p(out), 1 if out_nhwc else 0,
                                                _p(sp.hi) if sp else None, _p(sp.lo) if sp else None,
                                                ld_split or 0, _p(ws), wsb, _stream()), 'hvr_roi_align_fwd_fast')
    else:
        check(_lib.lib().hvr_roi_align_fwd(_p(feat), int(feat_nhwc), _p(rois), n, B, C, H, W, ph, pw, float(spatial_scale),
                                           int(sample_num), _p(out), 1 if out_nhwc else 0,
                                           _p(sp.hi) if sp else None, _p(sp.lo) if sp else None,
                                           ld_split or 0, _p(ws), _stream()), 'hvr_roi_align_fwd')
    return (out, sp) if want_split else out


def nms(dets, iou_thr, strict_gt=True):
    """dets [n,5] fp32 CUDA -> kept original indices (int64, ascending).  One D2H read of the
    count at the end (the reference API returns a tensor of data-dependent length)."""
    _need_cuda(dets)
    dets = dets.contiguous().float()
    n = dets.shape[0]
    dev = dets.device
    if n == 0:
        return torch.zeros(0, dtype=torch.long, device=dev)
    L = _lib.lib()
    wsb = L.hvr_nms_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    keep = torch.empty(n, dtype=torch.long, device=dev)
    nk = torch.zeros(1, dtype=torch.int32, device=dev)
    check(L.hvr_nms(_p(dets), n, float(iou_thr), int(strict_gt), _p(keep), _p(nk), _p(ws), wsb, _stream()), 'hvr_nms')
    return keep[:int(nk.item())]


def rpn_proposals(cls, reg, ld_cls, ld_reg, T, H, W, A, base_anchors, stride, img_shape, nms_pre=6000, nms_post=300,
                  max_num=300, nms_thr=0.7, want_idx=False):
    """cls/reg: fp32 CUDA tensors whose data_ptr is element (t=0, cell=0, a=0); see hvr_rpn_proposals.
    Returns proposals [T,max_num,5], counts [T] int32 (device) (+ anchor indices)."""
    _need_cuda(cls, reg, base_anchors)
    L = _lib.lib()
    dev = cls.device
    n_anc = H * W * A
    wsb = L.hvr_rpn_workspace_bytes(T, n_anc, nms_pre)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    props = torch.empty((T, max_num, 5), dtype=torch.float32, device=dev)
    counts = torch.empty(T, dtype=torch.int32, device=dev)
    idx = torch.empty((T, max_num), dtype=torch.int32, device=dev) if want_idx else None
    check(L.hvr_rpn_proposals(_p(cls), ld_cls, _p(reg), ld_reg, T, H, W, A, _p(base_anchors), stride,
                              float(img_shape[0]), float(img_shape[1]), nms_pre, nms_post, max_num, float(nms_thr),
                              _p(props), _p(counts), _p(idx), _p(ws), wsb, _stream()), 'hvr_rpn_proposals')
    return (props, counts, idx) if want_idx else (props, counts)


def det_postprocess(rois, cls, reg, img_shape, scale_factor=1.0, rescale=False, stds=(0.1, 0.1, 0.2, 0.2),
                    score_thr=0.001, iou_thr=0.3, max_per_img=300, n_cls=None):
    """rois [n,5], cls [n,>=n_cls] fp32, reg [n,>=4] fp32 (row-strided views allowed).
    Returns dets [max_per_img,5], labels [max_per_img] int64, n_dets [1] int32 (all device)."""
    _need_cuda(rois, cls, reg)
    L = _lib.lib()
    dev = rois.device
    n = rois.shape[0]
    n_cls = n_cls or cls.shape[1]
    rois = rois.contiguous().float()
    assert cls.stride(1) == 1 and reg.stride(1) == 1
    wsb = L.hvr_det_workspace_bytes(n, n_cls)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    dets = torch.zeros((max_per_img, 5), dtype=torch.float32, device=dev)
    labels = torch.zeros(max_per_img, dtype=torch.long, device=dev)
    nd = torch.zeros(1, dtype=torch.int32, device=dev)
    stds_c = (ctypes.c_float * 4)(*stds)
    check(L.hvr_det_postprocess(_p(rois), _p(cls), cls.stride(0), _p(reg), reg.stride(0), n, n_cls, stds_c,
                                float(img_shape[0]), float(img_shape[1]), float(scale_factor), int(rescale),
                                float(score_thr), float(iou_thr), int(max_per_img), _p(dets), _p(labels), _p(nd),
                                _p(ws), wsb, _stream()), 'hvr_det_postprocess')
    return dets, labels, nd


def det_postprocess_batched(rois, cls, reg, G, img_shape, scale_factor=1.0, rescale=False, stds=(0.1, 0.1, 0.2, 0.2),
                            score_thr=0.001, iou_thr=0.3, max_per_img=300, n_cls=None, n_valid=None, want_idx=False,
                            out=None):
    """G problems of n rois each in one launch per stage: rois [G*n,5], cls [G*n,>=n_cls], reg [G*n,>=4]
    (row-strided views allowed; problem g = rows [g*n, (g+1)*n)).  Returns dets [G,max_per_img,5],
    labels [G,max_per_img] int64, n_dets [G] int32; per problem bit-identical to det_postprocess.
    n_valid: optional device int32 [G] - only the first n_valid[g] rows of problem g are proposals.
    want_idx: also return roi_idx int32 [G,max_per_img] (the row every detection came from).
    out: optional preallocated (dets, labels, n_dets) - e.g. views of one packed result buffer."""
    _need_cuda(rois, cls, reg)
    L = _lib.lib()
    dev = rois.device
    assert rois.shape[0] % G == 0
    n = rois.shape[0] // G
    n_cls = n_cls or cls.shape[1]
    rois = rois.contiguous().float()
    assert cls.stride(1) == 1 and reg.stride(1) == 1
    wsb = L.hvr_det_batched_workspace_bytes(G, n, n_cls)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    if out is None:
        dets = torch.zeros((G, max_per_img, 5), dtype=torch.float32, device=dev)
        labels = torch.zeros((G, max_per_img), dtype=torch.long, device=dev)
        nd = torch.zeros(G, dtype=torch.int32, device=dev)
    else:
        dets, labels, nd = out
        assert dets.is_contiguous() and labels.is_contiguous() and nd.is_contiguous()
        assert dets.numel() == G * max_per_img * 5 and labels.dtype == torch.long and nd.dtype == torch.int32
    idx = torch.zeros((G, max_per_img), dtype=torch.int32, device=dev) if want_idx else None
    stds_c = (ctypes.c_float * 4)(*stds)
    check(L.hvr_det_postprocess_batched_ex(_p(rois), _p(cls), cls.stride(0), _p(reg), reg.stride(0), G, n, n_cls,
                                           stds_c, float(img_shape[0]), float(img_shape[1]), float(scale_factor),
                                           int(rescale), float(score_thr), float(iou_thr), int(max_per_img),
                                           _p(n_valid), _p(dets), _p(labels), _p(nd), _p(idx), _p(ws), wsb, _stream()),
          'hvr_det_postprocess_batched_ex')
    return (dets, labels, nd, idx) if want_idx else (dets, labels, nd)


def softmax_rows_split(S, cols, ld_p=None, seg_counts=None, slot=0, rows_per_problem=0):
    """S fp32 [rows, ld_s] -> Split P [rows, ld_p], softmax over the first `cols` columns.
    seg_counts (device int32 [n_problems, n_segs]): key mask for ragged proposal sets - of every block of `slot`
    columns only the first seg_counts[row // rows_per_problem][block] take part (hvr_softmax_rows_split_masked)."""
    _need_cuda(S)
    rows = S.shape[0]
    ld_p = ld_p or round_up(cols, 64)
    P = Split.empty((rows, ld_p), S.device)
    if seg_counts is None:
        check(_lib.lib().hvr_softmax_rows_split(_p(S), rows, cols, S.stride(0), _p(P.hi), _p(P.lo), ld_p, _stream()),
              'hvr_softmax_rows_split')
    else:
        assert seg_counts.dtype == torch.int32 and seg_counts.is_contiguous() and seg_counts.dim() == 2
        check(_lib.lib().hvr_softmax_rows_split_masked(_p(S), rows, cols, S.stride(0), _p(P.hi), _p(P.lo), ld_p,
                                                       _p(seg_counts), seg_counts.shape[1], int(slot),
                                                       int(rows_per_problem), _stream()),
              'hvr_softmax_rows_split_masked')
    return P


def window_rois(props, counts, perm, V, T, key_dim, n_segs=None, pad=True):
    """props [F,P,5], counts [F] int32 (hvr_rpn_proposals outputs), perm int64 [V*T] or None -> (rois [V*Npad,5],
    rois_key [V*P,5], seg_counts int32 [V,n_segs], key_counts int32 [V]); see hvr_window_rois.  pad=False: Npad = T*P
    (heads without attention need no 64-row alignment of a video's block)."""
    _need_cuda(props, counts)
    P = props.shape[1]
    Npad = round_up(T * P, 64) if pad else T * P
    n_segs = n_segs or T
    dev = props.device
    rois = torch.empty((V * Npad, 5), dtype=torch.float32, device=dev)
    rois_key = torch.empty((V * P, 5), dtype=torch.float32, device=dev)
    seg = torch.empty((V, n_segs), dtype=torch.int32, device=dev)      # entries >= T are zeroed by the kernel
    kc = torch.empty(V, dtype=torch.int32, device=dev)
    assert props.is_contiguous() and counts.dtype == torch.int32 and (perm is None or perm.dtype == torch.int64)
    check(_lib.lib().hvr_window_rois(_p(props), _p(counts), _p(perm), V, T, P, key_dim, Npad, _p(rois), _p(rois_key),
                                     _p(seg), n_segs, _p(kc), _stream()), 'hvr_window_rois')
    return rois, rois_key, seg, kc


def gather_rows(src, dst, n_problems, n_rows, idx=None, src_rpp=0, src_row0=0, dst_rpp=0, dst_row0=0, cols=None):
    """dst[p*dst_rpp + dst_row0 + j] = src[idx[p*n_rows + j]] (or src[p*src_rpp + src_row0 + j] when idx is None);
    Split matrices, negative index -> zero row.  See hvr_gather_rows_split."""
    cols = cols or src.shape[1]
    assert idx is None or (idx.dtype == torch.int32 and idx.is_contiguous())
    check(_lib.lib().hvr_gather_rows_split(_p(src.hi), _p(src.lo), src.hi.stride(0), _p(idx), src_rpp, src_row0,
                                           _p(dst.hi), _p(dst.lo), dst.hi.stride(0), n_problems, n_rows, dst_rpp,
                                           dst_row0, cols, _stream()), 'hvr_gather_rows_split')
    return dst


def support_index(sel, pool_counts, counts_rank_stride, vpr, rank_stride_rows, P, T, seg_counts):
    """sel int64 [V,S] global key-frame indices -> idx int32 [V, S*P] (rows of the gathered pool); fills the support
    blocks [T, T+S) of seg_counts [V, n_segs] in place.  See hvr_support_index."""
    V, S = sel.shape
    idx = torch.empty((V, S * P), dtype=torch.int32, device=sel.device)
    assert sel.dtype == torch.int64 and sel.is_contiguous() and pool_counts.dtype == torch.int32
    check(_lib.lib().hvr_support_index(_p(sel), _p(pool_counts), int(counts_rank_stride), int(vpr),
                                       int(rank_stride_rows), V, S, P, T, _p(idx), _p(seg_counts),
                                       seg_counts.shape[1], _stream()), 'hvr_support_index')
    return idx


def video_descriptor(c5_nhwc, n_videos):
    """c5_nhwc fp32 [n_videos*T, h, w, C] (the shared head's output, frames of a video contiguous) ->
    [n_videos, C]: max over a video's frames of the per-frame spatial mean (hnmb_rcnn.py:78-81)."""
    _need_cuda(c5_nhwc)
    c5_nhwc = c5_nhwc.contiguous()
    B, H, W, C = c5_nhwc.shape
    assert n_videos > 0 and B % n_videos == 0
    T = B // n_videos
    L = _lib.lib()
    nb = L.hvr_video_descriptor_workspace_bytes(n_videos, T, C)
    ws = torch.empty(nb, dtype=torch.uint8, device=c5_nhwc.device)
    desc = torch.empty((n_videos, C), dtype=torch.float32, device=c5_nhwc.device)
    check(L.hvr_video_descriptor(_p(c5_nhwc), n_videos, T, H * W, C, _p(desc), _p(ws), nb, _stream()),
          'hvr_video_descriptor')
    return desc


def support_select(desc, g0, n_local, n_support, want_weights=False):
    """desc fp32 [G, C] of all videos -> int64 [n_local, n_support]: for the videos g0 .. g0+n_local-1 the
    n_support other videos with the largest softmax similarity (hnmb_rcnn.py:85-88), ties to the lower
    index, -1 where fewer than n_support other videos exist.  Optionally the weights [n_local, G]."""
    _need_cuda(desc)
    desc = desc.contiguous().float()
    G, C = desc.shape
    idx = torch.full((n_local, n_support), -1, dtype=torch.int64, device=desc.device)
    w = torch.zeros((n_local, G), dtype=torch.float32, device=desc.device) if want_weights else None
    check(_lib.lib().hvr_support_select(_p(desc), G, C, g0, n_local, n_support, _p(idx), _p(w), _stream()),
          'hvr_support_select')
    return (idx, w) if want_weights else idx


def im2col_stem(img):
    img = img.contiguous().float()
    B, C, H, W = img.shape
    assert C == 3
    oh, ow = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    s = Split.empty((B, oh, ow, 192), img.device)
    check(_lib.lib().hvr_im2col_stem(_p(img), B, H, W, _p(s.hi), _p(s.lo), oh, ow, _stream()), 'hvr_im2col_stem')
    return s


def maxpool3x3s2(s):
    B, H, W, C = s.shape
    oh, ow = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    o = Split.empty((B, oh, ow, C), s.hi.device)
    check(_lib.lib().hvr_maxpool3x3s2_split(_p(s.hi), _p(s.lo), B, H, W, C, _p(o.hi), _p(o.lo), oh, ow, _stream()),
          'hvr_maxpool3x3s2_split')
    return o
