"""Functional forward of the hot path on the C-ABI kernels.

Weights come in as a flat state_dict with the REFERENCE's parameter names (SURVEY.md
section 5), are folded / re-laid-out / split once by the ``pack_*`` functions, and every
forward below is a sequence of libhvr_b200.so calls on the current stream.  Layout inside:
NHWC split-bf16 activations, K-major split weights [Cout, kh*kw*Cin].

Reference functions mirrored (file:line under /root/reference/mmdet):
  trunk_forward      models/backbones/resnet.py:222-257,522-533
  c5_forward         models/shared_heads/res_layer.py:67-74
  rpn_forward        models/anchor_heads/rpn_head.py:30-35
  relation           models/bbox_heads/hrnmp_bbox_head.py:216-355
  hrnmp_forward_test models/bbox_heads/hrnmp_bbox_head.py:800-909
  selsa_forward      models/bbox_heads/selsa_bbox_head.py:203-261
  shared_fc_forward  models/bbox_heads/convfc_bbox_head.py:126-167
"""
import math

import torch

from . import ops
from .ops import Split, round_up

BN_EPS = 1e-5
R101_BLOCKS = (3, 4, 23, 3)


# ----------------------------------------------------------------------------------------
# weight packing (host, once)
# ----------------------------------------------------------------------------------------

def _bn_fold(sd, name):
    g, b = sd[name + '.weight'].double(), sd[name + '.bias'].double()
    m, v = sd[name + '.running_mean'].double(), sd[name + '.running_var'].double()
    scale = g / torch.sqrt(v + BN_EPS)
    return scale, b - m * scale


def pack_matrix(w2d, device, pad_rows=64, pad_cols=64):
    """fp32/fp64 [N, K] -> Split [round_up(N), round_up(K)] on device (zero padded)."""
    n, k = w2d.shape
    W = torch.zeros((round_up(n, pad_rows), round_up(k, pad_cols)), dtype=torch.float32)
    W[:n, :k] = w2d.float()
    W = W.to(device)
    return ops.split(W)


def pack_conv(w, scale, device):
    """[Cout, Cin, kh, kw] (+ per-Cout BN scale) -> K-major [Cout, (r*kw+s)*Cin + c]."""
    w = w.double()
    if scale is not None:
        w = w * scale.view(-1, 1, 1, 1)
    cout = w.shape[0]
    return pack_matrix(w.permute(0, 2, 3, 1).reshape(cout, -1), device)


class ConvP:
    """One packed conv(+BN)(+bias) layer.  cin2 > 0: the weights carry a second K segment of cin2
    columns that multiplies a second input (the block's 1x1 downsample branch, _pack_conv3_down)."""
    __slots__ = ('w', 'bias', 'n', 'k', 'cin', 'dilation', 'cin2')

    def __init__(self, w, bias, n, k, cin, dilation=1, cin2=0):
        self.w, self.bias, self.n, self.k, self.cin, self.dilation, self.cin2 = w, bias, n, k, cin, dilation, cin2


def _pack_conv_bn(sd, conv, bn, device, dilation=1, bias_name=None):
    w = sd[conv + '.weight']
    scale = shift = None
    if bn is not None:
        scale, shift = _bn_fold(sd, bn)
    bias = shift
    if bias_name is not None:
        bias = sd[bias_name].double() if bias is None else bias + sd[bias_name].double() * scale
    n = w.shape[0]
    b = None
    if bias is not None:
        b = torch.zeros(round_up(n, 64), dtype=torch.float32)
        b[:n] = bias.float()
        b = b.to(device)
    return ConvP(pack_conv(w, scale, device), b, n, w.shape[2], w.shape[1], dilation)


# bn3(conv3(o)) + bn_d(conv_d(x)) of a block with a downsample branch (resnet.py:243-255) as ONE
# contraction over K = [conv3 channels | downsample channels] (hvr_igemm's second A operand): the
# 4x-wide identity tensor is neither written nor read back.  Set False for the two-GEMM evaluation.
FUSE_DOWNSAMPLE = True


def _pack_conv3_down(sd, q, device):
    w3, wd = sd[q + 'conv3.weight'].double(), sd[q + 'downsample.0.weight'].double()
    s3, b3 = _bn_fold(sd, q + 'bn3')
    sdn, bdn = _bn_fold(sd, q + 'downsample.1')
    n, c3, cd = w3.shape[0], w3.shape[1], wd.shape[1]
    assert c3 % 64 == 0 and w3.shape[2] == 1 and wd.shape[2] == 1
    w = torch.cat([(w3 * s3.view(-1, 1, 1, 1)).reshape(n, c3), (wd * sdn.view(-1, 1, 1, 1)).reshape(n, cd)], 1)
    b = torch.zeros(round_up(n, 64), dtype=torch.float32)
    b[:n] = (b3 + bdn).float()
    return ConvP(pack_matrix(w, device), b.to(device), n, 1, c3, 1, cin2=cd)


def _pack_res_layer(sd, p, blocks, dilation, device):
    out = []
    for i in range(blocks):
        q = '%s%d.' % (p, i)
        blk = dict(conv1=_pack_conv_bn(sd, q + 'conv1', q + 'bn1', device),
                   conv2=_pack_conv_bn(sd, q + 'conv2', q + 'bn2', device, dilation),
                   conv3=_pack_conv_bn(sd, q + 'conv3', q + 'bn3', device))
        if (q + 'downsample.0.weight') in sd:
            blk['down'] = _pack_conv_bn(sd, q + 'downsample.0', q + 'downsample.1', device)
            if sd[q + 'conv3.weight'].shape[1] % 64 == 0:
                blk['conv3d'] = _pack_conv3_down(sd, q, device)
        out.append(blk)
    return out


def pack_trunk(sd, device, prefix='backbone.', strides=(1, 2, 2), dilations=(1, 1, 1)):
    w = sd[prefix + 'conv1.weight'].double()
    scale, shift = _bn_fold(sd, prefix + 'bn1')
    w = (w * scale.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).reshape(64, 147)      # k = (r*7+s)*3 + c
    stem = ConvP(pack_matrix(w, device, 64, 192), shift.float().to(device), 64, 7, 3)
    layers = [_pack_res_layer(sd, '%slayer%d.' % (prefix, i + 1), R101_BLOCKS[i], dilations[i], device)
              for i in range(len(strides))]
    return dict(stem=stem, layers=layers, strides=tuple(strides))


def pack_c5(sd, device, prefix='shared_head.', stride=1, dilation=2):
    p = dict(layer4=_pack_res_layer(sd, prefix + 'layer4.', R101_BLOCKS[3], dilation, device), stride=stride)
    if (prefix + 'new_layer_1.conv.weight') in sd:
        p['new1'] = _pack_conv_bn(sd, prefix + 'new_layer_1.conv', None, device,
                                  bias_name=prefix + 'new_layer_1.conv.bias')
    return p


def pack_rpn(sd, device, prefix='rpn_head.'):
    conv = _pack_conv_bn(sd, prefix + 'rpn_conv', None, device, bias_name=prefix + 'rpn_conv.bias')
    wc, wr = sd[prefix + 'rpn_cls.weight'], sd[prefix + 'rpn_reg.weight']
    A = wc.shape[0]
    w = torch.cat([wc, wr], 0).reshape(A * 5, -1)                                # one GEMM: [cls | reg]
    b = torch.zeros(round_up(A * 5, 64), dtype=torch.float32)
    b[:A * 5] = torch.cat([sd[prefix + 'rpn_cls.bias'], sd[prefix + 'rpn_reg.bias']])
    head = ConvP(pack_matrix(w, device), b.to(device), A * 5, 1, w.shape[1])
    return dict(conv=conv, head=head, A=A)


class LinP:
    __slots__ = ('w', 'bias', 'n', 'k')

    def __init__(self, w, bias, n, k):
        self.w, self.bias, self.n, self.k = w, bias, n, k


def pack_linear(weight, bias, device, col_perm=None):
    w = weight.reshape(weight.shape[0], -1)
    if col_perm is not None:
        w = w[:, col_perm]
    n, k = w.shape
    b = torch.zeros(round_up(n, 64), dtype=torch.float32)
    if bias is not None:
        b[:n] = bias
    return LinP(pack_matrix(w, device), b.to(device), n, k)


def nhwc_perm(channels, size):
    """Column permutation taking flatten(C,h,w) weights to flatten(h,w,C) inputs."""
    idx = torch.arange(channels * size * size).view(channels, size, size)
    return idx.permute(1, 2, 0).reshape(-1)


def pack_head(sd, device, kind='hrnmp', prefix='bbox_head.', roi_channels=256, roi_size=7):
    """kind: 'hrnmp' (4 stages) | 'selsa' (2 stages) | 'shared_fc'."""
    perm = nhwc_perm(roi_channels, roi_size)
    P = dict(kind=kind)

    def lin(name, col_perm=None):
        return pack_linear(sd[prefix + name + '.weight'], sd[prefix + name + '.bias'], device, col_perm)

    if kind == 'shared_fc':
        P['fc1'] = lin('shared_fcs.0', perm)
        P['fc2'] = lin('shared_fcs.1')
    else:
        stages = 4 if kind == 'hrnmp' else 2
        for k in range(1, stages + 1):
            P['fc%d' % k] = lin('fc_new_%d' % k, perm if k == 1 else None)
            s = 'selsa_%d.' % k
            P['q%d' % k] = lin(s + 'q_data_fc_%d' % k)
            P['k%d' % k] = lin(s + 'k_data_fc_%d' % k)
            # stages whose queries are all rows (1, 3): q_data_fc_k and k_data_fc_k read the same X
            # (hrnmp_bbox_head.py:282-291) -> ONE GEMM with N = 2048, columns [Q | K]; X is streamed once
            if k % 2 == 1:
                P['qk%d' % k] = pack_linear(
                    torch.cat([sd[prefix + s + 'q_data_fc_%d.weight' % k], sd[prefix + s + 'k_data_fc_%d.weight' % k]], 0),
                    torch.cat([sd[prefix + s + 'q_data_fc_%d.bias' % k], sd[prefix + s + 'k_data_fc_%d.bias' % k]], 0), device)
            P['o%d' % k] = lin(s + 'linear_out_%d' % k)
    # fc_cls and fc_reg share their input: one GEMM, columns [cls | reg]
    def clsreg(c, r):
        w = torch.cat([sd[prefix + c + '.weight'], sd[prefix + r + '.weight']], 0)
        b = torch.cat([sd[prefix + c + '.bias'], sd[prefix + r + '.bias']], 0)
        return pack_linear(w, b, device)
    P['n_cls'] = sd[prefix + 'fc_cls.weight'].shape[0]
    P['n_reg'] = sd[prefix + 'fc_reg.weight'].shape[0]
    P['out1'] = clsreg('fc_cls', 'fc_reg')
    if kind == 'hrnmp':
        P['out2'] = clsreg('fc_cls_2', 'fc_reg_2')
    return P


# ----------------------------------------------------------------------------------------
# conv / linear execution
# ----------------------------------------------------------------------------------------

def _taps(k, dil):
    r = k // 2
    return tuple(((s - r) * dil, (t - r) * dil) for t in range(k) for s in range(k))   # (dx, dy), r-major


def conv(a, cp, stride=1, relu=False, res=None, want_split=True, want_f32=False, passes=3, check_kernel=False,
         a2=None, a2_stride=1):
    """a: Split NHWC [B,H,W,C].  Returns (Split NHWC or None, fp32 NHWC or None).
    a2: second input Split NHWC [B,H2,W2,cp.cin2] sampled at (y*a2_stride, x*a2_stride) of the output
    pixel (the 1x1 downsample branch folded into the weights, _pack_conv3_down)."""
    B, H, W, C = a.shape
    assert C == cp.cin, (C, cp.cin)
    dev = a.hi.device
    sb, sh, sw = a.hi.stride(0), a.hi.stride(1), a.hi.stride(2)
    if stride == 1:
        Ho, Wo = H, W
        view = (C, W, H, B, sw, sh, sb)
    else:
        assert cp.k == 1, 'strided convs on this path are 1x1 (caffe-style bottleneck)'
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        view = (C, Wo, Ho, B, sw * stride, sh * stride, sb)
    n = cp.w.shape[0]          # physical (64-padded) rows; padded outputs are exactly bias-free zeros
    out = Split.empty((B, Ho, Wo, n), dev) if want_split else None
    of = torch.empty((B, Ho, Wo, round_up(n, 4)), dtype=torch.float32, device=dev) if want_f32 else None
    rows = B * Ho * Wo
    a2_view = None
    if a2 is not None:
        B2, H2, W2, C2 = a2.shape
        assert C2 == cp.cin2 and B2 == B and (H2 - 1) // a2_stride + 1 == Ho and (W2 - 1) // a2_stride + 1 == Wo
        a2_view = (C2, Wo, Ho, B, a2.hi.stride(2) * a2_stride, a2.hi.stride(1) * a2_stride, a2.hi.stride(0))
    g = ops.igemm_desc(a, cp.w, n, taps=_taps(cp.k, cp.dilation), a_view=view, out_whb=(Wo, Ho, B),
                       bias=cp.bias, relu=relu, a2=a2, a2_view=a2_view,
                       res=Split(res.hi.view(rows, -1), res.lo.view(rows, -1)) if res is not None else None,
                       out=Split(out.hi.view(rows, n), out.lo.view(rows, n)) if out is not None else None,
                       out_f32=of.view(rows, -1) if of is not None else None, passes=passes)
    ops.igemm_run(g, check_kernel)
    return out, of


def bottleneck(x, blk, stride, **kw):
    """resnet.py:222-257, caffe style: stride on conv1 (and on the downsample)."""
    o, _ = conv(x, blk['conv1'], stride=stride, relu=True, **kw)
    o, _ = conv(o, blk['conv2'], relu=True, **kw)
    if FUSE_DOWNSAMPLE and 'conv3d' in blk:
        o, _ = conv(o, blk['conv3d'], relu=True, a2=x, a2_stride=stride, **kw)
        return o
    idt = x
    if 'down' in blk:
        idt, _ = conv(x, blk['down'], stride=stride, **kw)
    o, _ = conv(o, blk['conv3'], relu=True, res=idt, **kw)
    return o


def res_layer(x, blocks, stride, **kw):
    for i, blk in enumerate(blocks):
        x = bottleneck(x, blk, stride if i == 0 else 1, **kw)
    return x


def trunk_forward(P, img, **kw):
    """img fp32 NCHW [B,3,H,W] -> C4 Split NHWC [B,H/16,W/16,1024]."""
    col = ops.im2col_stem(img)                                   # [B,oh,ow,192]
    x, _ = _stem(P, col, **kw)
    x = ops.maxpool3x3s2(x)
    for blocks, s in zip(P['layers'], P['strides']):
        x = res_layer(x, blocks, s, **kw)
    return x


def _stem(P, col, **kw):
    st = P['stem']
    cp = ConvP(st.w, st.bias, 64, 1, 192)
    return conv(col, cp, relu=True, **kw)


def c5_forward(P, c4, **kw):
    """C4 Split NHWC -> fp32 NHWC [T,h,w,256] (RoIAlign input)."""
    x = res_layer(c4, P['layer4'], P['stride'], **kw)
    if 'new1' in P:
        _, f = conv(x, P['new1'], relu=True, want_split=False, want_f32=True, **kw)
        return f
    return ops.merge(x)


def rpn_forward(P, c4, **kw):
    """-> fp32 [T,h,w,64]: columns [0,A) cls logits, [A,5A) deltas (a*4+d)."""
    x, _ = conv(c4, P['conv'], relu=True, **kw)
    _, o = conv(x, P['head'], want_split=False, want_f32=True, **kw)
    return o


def lin(a, lp, relu=False, res=None, want_split=True, want_f32=False, want_T=False, **kw):
    """Linear layer.  want_T: also X^T (Split [n, round_up(M, 64)]) - produced by a separate transpose
    of the split output (two 16-byte-vectorised passes over 4 B/element) rather than by the GEMM's
    transposed-store epilogue, which made the fc_new_k launches 2x slower (2-byte scattered stores)."""
    n = lp.w.shape[0]
    out, of, _ = ops.linear(a, lp.w, n, bias=lp.bias, relu=relu, res=res, want_split=want_split or want_T,
                            want_f32=want_f32, **kw)
    oT = ops.transpose_split(out, n) if want_T else None
    return (out if want_split else None), of, oT


# q_data_fc_k and k_data_fc_k of an all-row stage as one N = 2048 GEMM (False: two GEMMs, the same bits)
FUSE_QK = True


def _qk(P, idx, X, **kw):
    """-> (Q, K) as column halves of one [M, 2048] product (row pitch 2048)."""
    lp = P['qk%d' % idx]
    D = lp.n // 2
    QK, _, _ = lin(X, lp, **kw)
    return Split(QK.hi[:, :D], QK.lo[:, :D]), Split(QK.hi[:, D:2 * D], QK.lo[:, D:2 * D])


def relation(P, idx, X, XT, q_range=None, res=None, relu=True, extra=None, **kw):
    """SELSA relation block idx on X Split [N,D] (XT = X^T Split [D, ld>=N]).
    Returns relu(res + NL(X)) (Split [Nq, D]).  hrnmp_bbox_head.py:216-355.
    extra: optional (Xs Split [M,D]) support rows appended to the key/value set (stage 4)."""
    N, D = X.shape
    if extra is not None:
        Xk = Split(torch.cat([X.hi, extra.hi], 0), torch.cat([X.lo, extra.lo], 0))
        Nk = Xk.shape[0]
        XkT = Split.zeros((D, round_up(Nk, 64)), X.hi.device)
        XkT.hi[:, :Nk] = Xk.hi.t()
        XkT.lo[:, :Nk] = Xk.lo.t()
    else:
        Xk, Nk, XkT = X, N, XT
    if q_range is None and extra is None and FUSE_QK and ('qk%d' % idx) in P:
        Q, K = _qk(P, idx, X, **kw)
    else:
        Xq = X if q_range is None else X[q_range[0]:q_range[0] + q_range[1]]
        Q, _, _ = lin(Xq, P['q%d' % idx], **kw)
        K, _, _ = lin(Xk, P['k%d' % idx], **kw)
    # S = Q K^T / sqrt(D)  (1/32 for D=1024: exact power of two)
    _, S, _ = ops.linear(Q, K, Nk, alpha=1.0 / math.sqrt(float(D)), want_split=False, want_f32=True, **kw)
    Pm = ops.softmax_rows_split(S, Nk)
    # O = P V, V = un-projected rows (conv_g False): right operand is X^T [D, Nk]
    O, _, _ = ops.linear(Pm, XkT, D, **kw)
    out, _, _ = lin(O, P['o%d' % idx], relu=relu, res=res, **kw)
    return out


# ----------------------------------------------------------------------------------------
# Batched-over-videos head: V windows of N = T*slot rows each, laid out [V*Npad, D] (Npad = N rounded up to
# 64; pad rows are finite garbage that never reaches a result).  Every row-wise GEMM (fc_new_k, Q/K/out
# projections, cls|reg) runs ONCE over all videos (M = V*Npad fills the machine with 256-row tile pairs); QK^T
# and P.V are one batched launch each (ops.bmm: a per-video B matrix), the softmax one launch over all V*Nq
# rows.  Per-row arithmetic and the per-video attention operands are those of the single-window functions
# above, so the results are bit-identical to them.
#
# Ragged proposal sets (hnmb_rcnn.py:582-599 uses the ACTUAL per-frame counts): every frame keeps a fixed block
# of `slot` rows; `mask` = KeyMask(seg_counts int32 [V, n_segs] on the device, slot) tells the softmax which keys
# of each block are proposals.  Rows behind a frame's count are finite garbage: as keys they get probability 0,
# as queries they are never read.  With every count == slot the mask changes no bit.
# ----------------------------------------------------------------------------------------
class KeyMask:
    __slots__ = ('seg', 'slot')

    def __init__(self, seg, slot):
        self.seg, self.slot = seg, slot


def key_rows(X, V, Npad, s, n):
    """Rows [s, s+n) of every video's block as one Split [V*n, D] (one gather launch)."""
    out = Split.empty((V * n, X.shape[1]), X.hi.device)
    return ops.gather_rows(X, out, V, n, src_rpp=Npad, src_row0=s, dst_rpp=n)


def attention_batched(Q, K, XT, V, nk, nk_pad, mask=None):
    """softmax(Q_v K_v^T / sqrt(D)) X_v for V videos.  Q Split [V*nq, D]; K Split [V*nk_pad, D] (video-major
    blocks, first nk rows of a block are keys); XT Split [D, V*nk_pad] (values = un-projected rows, conv_g
    False).  One launch each: QK^T (hvr_igemm with a per-image B matrix), row softmax, P.V."""
    D = Q.shape[1]
    nq = Q.shape[0] // V
    _, S = ops.bmm(Q, K, V, nk, nk_pad * K.hi.stride(0), alpha=1.0 / math.sqrt(float(D)), want_split=False,
                   want_f32=True)
    Pm = ops.softmax_rows_split(S, nk, ld_p=nk_pad, seg_counts=mask.seg if mask is not None else None,
                                slot=mask.slot if mask is not None else 0, rows_per_problem=nq)
    O, _ = ops.bmm(Pm, XT, V, D, nk_pad)
    return O


def relation_batched(P, idx, X, XT, V, N, Npad, q_range=None, Xkey=None, res=None, relu=True, mask=None, **kw):
    """Relation block idx for V videos.  q_range None: all rows are queries (res = X); else the key rows
    (Xkey = key_rows(X, ...), also the residual)."""
    if q_range is None and FUSE_QK and ('qk%d' % idx) in P:
        Q, K = _qk(P, idx, X, **kw)
    else:
        if q_range is None:
            Xq = X
        else:
            Xq = Xkey if Xkey is not None else key_rows(X, V, Npad, q_range[0], q_range[1])
        Q, _, _ = lin(Xq, P['q%d' % idx], **kw)
        K, _, _ = lin(X, P['k%d' % idx], **kw)
    O = attention_batched(Q, K, XT, V, N, Npad, mask)
    out, _, _ = lin(O, P['o%d' % idx], relu=relu, res=res, **kw)
    return out


def hrnmp_stage123_batched(P, rows, V, N, Npad, start, length, mask=None, f1=None, f1T=None, **kw):
    """Stages 1-3 + fc_new_4 for V windows at once.  rows Split [V*Npad, 12544] (or f1 / f1T: the cached
    fc_new_1 rows of the streaming scheduler).  Returns (out1 fp32 [V*length, 64], f4 Split [V*Npad, D],
    f4^T Split [D, V*Npad])."""
    s, n = start, length
    if f1 is None:
        f1, _, f1T = lin(rows, P['fc1'], want_T=True, **kw)
    a1 = relation_batched(P, 1, f1, f1T, V, N, Npad, res=f1, mask=mask, **kw)
    f2, _, f2T = lin(a1, P['fc2'], want_T=True, **kw)
    f2k = key_rows(f2, V, Npad, s, n)
    a2k = relation_batched(P, 2, f2, f2T, V, N, Npad, q_range=(s, n), Xkey=f2k, res=f2k, mask=mask, **kw)
    _, out1, _ = lin(a2k, P['out1'], want_split=False, want_f32=True, **kw)
    # input of stage 3: the fc_new_1 rows with the key rows replaced by stage 2's output (hrnmp_bbox_head.py:865-868)
    x3 = Split.empty(tuple(f1.shape), f1.hi.device)
    ops.gather_rows(f1, x3, 1, V * Npad)
    ops.gather_rows(a2k, x3, V, n, src_rpp=n, dst_rpp=Npad, dst_row0=s)
    f3, _, f3T = lin(x3, P['fc3'], want_T=True, **kw)
    a3 = relation_batched(P, 3, f3, f3T, V, N, Npad, res=f3, mask=mask, **kw)
    f4, _, f4T = lin(a3, P['fc4'], want_T=True, **kw)
    return out1, f4, f4T


def hrnmp_stage4_batched(P, f4, f4T, V, N, Npad, start, length, mask=None, f4k=None, **kw):
    s, n = start, length
    if f4k is None:
        f4k = key_rows(f4, V, Npad, s, n)
    a4 = relation_batched(P, 4, f4, f4T, V, N, Npad, q_range=(s, n), Xkey=f4k, res=f4k, mask=mask, **kw)
    _, out2, _ = lin(a4, P['out2'], want_split=False, want_f32=True, **kw)
    return out2


def hrnmp_forward_batched(P, rows, V, N, Npad, start, length, mask=None, f1=None, f1T=None, **kw):
    """rows Split [V*Npad, 12544].  Returns fp32 (out1, out2) of shape [V*length, 64]."""
    out1, f4, f4T = hrnmp_stage123_batched(P, rows, V, N, Npad, start, length, mask=mask, f1=f1, f1T=f1T, **kw)
    return out1, hrnmp_stage4_batched(P, f4, f4T, V, N, Npad, start, length, mask=mask, **kw)


def hrnmp_stage4_inter_batched(P, f4, f4k, Kown, sup_rows, V, N, Npad, length, n_sup_rows, mask=None, **kw):
    """Stage 4 with inter-video support rows (hrnmp_bbox_head.py:740-795 at inference):
    the key / value set of video v = its own N window rows followed by its n_sup_rows support rows
    (sup_rows Split [V*n_sup_rows, D]: the post-fc_new_4 key rows of the chosen other key frames, already
    gathered).  Kown = k_data_fc_4(f4) [V*Npad, D] (computed while the exchange is in flight), f4k = the key rows
    (queries and residual).  mask: seg_counts with the support blocks filled in."""
    D = f4.shape[1]
    n = length
    nk = N + n_sup_rows
    nk_pad = round_up(nk, 64)
    dev = f4.hi.device
    Q, _, _ = lin(f4k, P['q4'], **kw)
    Ksup, _, _ = lin(sup_rows, P['k4'], **kw)
    # assemble per video [own N rows | support rows | zero pad] for the keys and the values
    Kall = Split.empty((V * nk_pad, D), dev)
    Xall = Split.empty((V * nk_pad, D), dev)
    for src_own, src_sup, dst in ((Kown, Ksup, Kall), (f4, sup_rows, Xall)):
        ops.gather_rows(src_own, dst, V, N, src_rpp=Npad, dst_rpp=nk_pad)
        ops.gather_rows(src_sup, dst, V, n_sup_rows, src_rpp=n_sup_rows, dst_rpp=nk_pad, dst_row0=N)
    if nk_pad > nk:   # zero the pad rows of the value set (P is exactly 0 there; 0 * garbage must stay 0)
        zero_idx = torch.full((V * (nk_pad - nk),), -1, dtype=torch.int32, device=dev)
        ops.gather_rows(f4, Xall, V, nk_pad - nk, idx=zero_idx, dst_rpp=nk_pad, dst_row0=nk)
    XallT = ops.transpose_split(Xall, D)
    O = attention_batched(Q, Kall, XallT, V, nk, nk_pad, mask)
    a4, _, _ = lin(O, P['o4'], relu=True, res=f4k, **kw)
    _, out2, _ = lin(a4, P['out2'], want_split=False, want_f32=True, **kw)
    return out2


def selsa_forward_batched(P, rows, V, N, Npad, start, length, mask=None, f1=None, f1T=None, **kw):
    s, n = start, length
    if f1 is None:
        f1, _, f1T = lin(rows, P['fc1'], want_T=True, **kw)
    a1 = relation_batched(P, 1, f1, f1T, V, N, Npad, res=f1, mask=mask, **kw)
    f2, _, f2T = lin(a1, P['fc2'], want_T=True, **kw)
    f2k = key_rows(f2, V, Npad, s, n)
    a2k = relation_batched(P, 2, f2, f2T, V, N, Npad, q_range=(s, n), Xkey=f2k, res=f2k, mask=mask, **kw)
    _, out1, _ = lin(a2k, P['out1'], want_split=False, want_f32=True, **kw)
    return out1


def head_fc1(P, roi_feats, **kw):
    """fc_new_1 on RoI rows (hrnmp_bbox_head.py:827-828): per-row, so a frame's rows can be computed
    once and reused by every window the frame appears in.  Returns (f1 Split [n,D], f1^T Split [D,ld])."""
    f1, _, f1T = lin(roi_feats, P['fc1'], want_T=True, **kw)
    return f1, f1T


def hrnmp_stage123(P, roi_feats, start, length, f1=None, f1T=None, **kw):
    """Stages 1-3 + fc_new_4 of forward_test (hrnmp_bbox_head.py:827-889).  Returns
    (out1 fp32 [len, 64] = branch [cls | reg], f4 Split [N, D], f4^T Split [D, ld])."""
    s, e = start, start + length
    if f1 is None:
        f1, f1T = head_fc1(P, roi_feats, **kw)
    a1 = relation(P, 1, f1, f1T, res=f1, **kw)
    f2, _, f2T = lin(a1, P['fc2'], want_T=True, **kw)
    # only the key rows of stage 2 are ever used (hrnmp_bbox_head.py:865-868): key-only queries
    a2k = relation(P, 2, f2, f2T, q_range=(s, length), res=f2[s:e], **kw)
    _, out1, _ = lin(a2k, P['out1'], want_split=False, want_f32=True, **kw)
    x3 = Split(torch.cat([f1.hi[:s], a2k.hi, f1.hi[e:]], 0), torch.cat([f1.lo[:s], a2k.lo, f1.lo[e:]], 0))
    f3, _, f3T = lin(x3, P['fc3'], want_T=True, **kw)
    a3 = relation(P, 3, f3, f3T, res=f3, **kw)
    f4, _, f4T = lin(a3, P['fc4'], want_T=True, **kw)
    return out1, f4, f4T


def hrnmp_stage4(P, f4, f4T, start, length, support=None, **kw):
    """Stage 4 (hrnmp_bbox_head.py:888-906): key-row queries over all rows of the window plus
    optional inter-video support rows (post-fc_new_4 key rows of other videos)."""
    s, e = start, start + length
    a4 = relation(P, 4, f4, f4T, q_range=(s, length), res=f4[s:e], extra=support, **kw)
    _, out2, _ = lin(a4, P['out2'], want_split=False, want_f32=True, **kw)
    return out2


def hrnmp_forward_test(P, roi_feats, start, length, support=None, f1=None, f1T=None, **kw):
    """roi_feats Split [N, 12544] (NHWC-flattened).  Returns fp32 (out1 [len, 64], out2 [len, 64]):
    columns [0,n_cls) class logits, [n_cls, n_cls+4) box deltas; plus f4[key] Split.
    f1/f1T: precomputed fc_new_1 rows of the window (streaming caches), else computed here."""
    out1, f4, f4T = hrnmp_stage123(P, roi_feats, start, length, f1=f1, f1T=f1T, **kw)
    out2 = hrnmp_stage4(P, f4, f4T, start, length, support, **kw)
    return out1, out2, f4[start:start + length]


def selsa_forward(P, roi_feats, start, length, f1=None, f1T=None, **kw):
    s, e = start, start + length
    if f1 is None:
        f1, f1T = head_fc1(P, roi_feats, **kw)
    a1 = relation(P, 1, f1, f1T, res=f1, **kw)
    f2, _, f2T = lin(a1, P['fc2'], want_T=True, **kw)
    a2k = relation(P, 2, f2, f2T, q_range=(s, length), res=f2[s:e], **kw)   # relu after the key slice
    _, out1, _ = lin(a2k, P['out1'], want_split=False, want_f32=True, **kw)
    return out1


def shared_fc_forward(P, roi_feats, **kw):
    x, _, _ = lin(roi_feats, P['fc1'], relu=True, **kw)
    x, _, _ = lin(x, P['fc2'], relu=True, **kw)
    _, out1, _ = lin(x, P['out1'], want_split=False, want_f32=True, **kw)
    return out1
