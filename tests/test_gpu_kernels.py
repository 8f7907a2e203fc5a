"""Kernel-level parity on the GPU: every C-ABI entry point against the CPU oracle
(oracle/) on the same seeded inputs.  Integer / index outputs are compared bit-exactly,
fp32 outputs within the tolerance written next to each assert."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ---------------------------------------------------------------------------------------
# split-bf16 storage
# ---------------------------------------------------------------------------------------
def test_split_merge_roundtrip(cuda):
    from hvrnet_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(100003, generator=g) * torch.logspace(-3, 3, 100003)).to(cuda)
    s = ops.split(x)
    y = ops.merge(s)
    # hi + lo carries >= 16 mantissa bits: relative error <= 2^-16 per element
    assert float(((y - x).abs() / x.abs().clamp_min(1e-30)).max()) < 2.0 ** -16
    assert torch.equal(s.hi, x.to(torch.bfloat16))


def test_layout_roundtrip(cuda):
    from hvrnet_b200 import ops
    x = torch.randn(2, 37, 5, 9, device=cuda)
    assert torch.equal(ops.nhwc_to_nchw(ops.nchw_to_nhwc(x)), x)
    s = ops.nchw_to_nhwc_split(x)
    assert s.hi.shape == (2, 5, 9, 37)
    y = ops.nhwc_split_to_nchw(s)
    assert _rel(y, x) < 2.0 ** -16


# ---------------------------------------------------------------------------------------
# implicit GEMM (tcgen05) against fp64 torch on the merged operands
# ---------------------------------------------------------------------------------------
def _ref_linear(a, w, n, bias, res, relu, alpha):
    A = a.float().double().cpu()
    W = w.float().double().cpu()[:n, :A.shape[1]]
    y = alpha * (A @ W.t())
    if bias is not None:
        y = y + bias.double().cpu()[:n]
    if res is not None:
        y = y + res.float().double().cpu()[:, :n]
    if relu:
        y = y.clamp_min(0)
    return y


@pytest.mark.parametrize('M,K,N,relu,use_res,use_bias', [
    (128, 64, 64, False, False, False),        # one tile, one K step
    (300, 1024, 1024, True, True, True),       # relation-head shape
    (257, 12544, 128, False, False, True),     # fc_new_1 K
    (200, 1024, 64, False, False, True),       # narrow N (fc_cls|fc_reg padded)
    (333, 448, 300, False, False, False),      # ragged M and N, 7 K steps
])
def test_igemm_linear(cuda, M, K, N, relu, use_res, use_bias):
    from hvrnet_b200 import ops
    g = torch.Generator().manual_seed(M + K + N)
    a = ops.split(torch.randn(M, K, generator=g).to(cuda))
    w = ops.split((torch.randn(ops.round_up(N, 64), K, generator=g) / math.sqrt(K)).to(cuda))
    bias = torch.randn(ops.round_up(N, 64), generator=g).to(cuda) if use_bias else None
    res = ops.split(torch.randn(M, ops.round_up(N, 8), generator=g).to(cuda)) if use_res else None
    out, of, oT = ops.linear(a, w, N, bias=bias, relu=relu, res=res, alpha=0.5, want_split=True, want_f32=True,
                             want_T=True)
    torch.cuda.synchronize()
    ref = _ref_linear(a, w, N, bias, res, relu, 0.5)
    got = of[:, :N].double().cpu()
    err = float((got - ref).abs().max() / ref.abs().max())
    # three-product split-bf16: dropped lo*lo and a_3 terms ~2^-16 per product, fp32 accumulation
    assert err < 3e-5, err
    assert _rel(ops.merge(Split_rows(out))[:, :N].double().cpu(), ref) < 3e-5
    assert _rel(ops.merge(Split_rows(oT))[:N, :M].t().double().cpu(), ref) < 3e-5
    # the SIMT cross-check kernel evaluates the same descriptor
    _, of2, _ = ops.linear(a, w, N, bias=bias, relu=relu, res=res, alpha=0.5, want_split=False, want_f32=True,
                           check_kernel=True)
    assert _rel(of2[:, :N].double().cpu(), ref) < 1e-5


@pytest.mark.parametrize('bn', [64, 128, 256, 512, 640])
def test_igemm_tile_widths(cuda, bn):
    """Every N-tile instantiation of the persistent kernel (512 / 640 = the cta_group::2 pair kernel with
    256- / 128-wide tiles) on a
    multi-tile, multi-wave problem (more tiles than SMs, so each CTA walks several tiles through
    both TMEM buffers)."""
    from hvrnet_b200 import _lib, ops
    g = torch.Generator().manual_seed(bn)
    M, K, N = 128 * 41 + 17, 320, 768
    a = ops.split(torch.randn(M, K, generator=g).to(cuda))
    w = ops.split((torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda))
    bias = torch.randn(N, generator=g).to(cuda)
    assert _lib.lib().hvr_debug_force_bn(bn) == 0
    try:
        _, of, _ = ops.linear(a, w, N, bias=bias, relu=True, want_split=False, want_f32=True)
        torch.cuda.synchronize()
    finally:
        _lib.lib().hvr_debug_force_bn(0)
    ref = _ref_linear(a, w, N, bias, None, True, 1.0)
    assert _rel(of.double().cpu(), ref) < 3e-5


@pytest.mark.parametrize('M,K,N', [(128 * 43, 320, 768), (100, 64, 128), (128 * 3 + 5, 1024, 4500)])
def test_igemm_cta_pair_odd_tiles(cuda, M, K, N):
    """cta_group::2 kernel: odd number of 128-row tiles (the peer CTA of the last pair runs on
    zero-filled rows), ragged N (last B half-tile out of range), residual + transposed outputs."""
    from hvrnet_b200 import _lib, ops
    g = torch.Generator().manual_seed(M + N)
    a = ops.split(torch.randn(M, K, generator=g).to(cuda))
    w = ops.split((torch.randn(ops.round_up(N, 64), K, generator=g) / math.sqrt(K)).to(cuda))
    bias = torch.randn(ops.round_up(N, 64), generator=g).to(cuda)
    res = ops.split(torch.randn(M, ops.round_up(N, 8), generator=g).to(cuda))
    assert _lib.lib().hvr_debug_force_bn(512) == 0
    try:
        out, of, oT = ops.linear(a, w, N, bias=bias, relu=False, res=res, want_split=True, want_f32=True, want_T=True)
        torch.cuda.synchronize()
    finally:
        _lib.lib().hvr_debug_force_bn(0)
    ref = _ref_linear(a, w, N, bias, res, False, 1.0)
    assert _rel(of[:, :N].double().cpu(), ref) < 3e-5
    assert _rel(ops.merge(Split_rows(out))[:, :N].double().cpu(), ref) < 3e-5
    assert _rel(ops.merge(Split_rows(oT))[:N, :M].t().double().cpu(), ref) < 3e-5


def test_igemm_cta_pair_conv(cuda):
    import torch.nn.functional as F
    from hvrnet_b200 import _lib, engine, ops
    g = torch.Generator().manual_seed(8)
    B, H, W, C, N, k, dil = 3, 38, 63, 128, 256, 3, 2
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(N, C, k, k, generator=g) / math.sqrt(C * k * k)
    xs = ops.nchw_to_nhwc_split(x.to(cuda))
    cp = engine.ConvP(engine.pack_conv(w, None, cuda), torch.zeros(N, device=cuda), N, k, C, dil)
    _lib.lib().hvr_debug_force_bn(512)
    try:
        out, _ = engine.conv(xs, cp, relu=True)
        torch.cuda.synchronize()
    finally:
        _lib.lib().hvr_debug_force_bn(0)
    xm = ops.nhwc_split_to_nchw(xs).double().cpu()
    wm = ops.merge(cp.w).double().cpu()[:N].view(N, k, k, C).permute(0, 3, 1, 2)
    ref = F.conv2d(xm, wm, padding=dil, dilation=dil).clamp_min(0)
    assert _rel(ops.nhwc_split_to_nchw(out).double().cpu(), ref) < 3e-5


@pytest.mark.parametrize('force', [0, 64, 128, 256, 512, 640])
def test_igemm_tma_epilogue_matches_per_row_epilogue(cuda, force):
    """The TMA epilogue (residual tiles loaded / split output tiles stored by TMA through swizzled
    shared memory) writes the same bits as the per-row register epilogue (flag 1024)."""
    from hvrnet_b200 import _lib, ops
    g = torch.Generator().manual_seed(77 + force)
    M, K, N = 128 * 9 + 50, 256, 600          # ragged M and N (N tail: 600 = 18*32 + 24)
    a = ops.split(torch.randn(M, K, generator=g).to(cuda))
    w = ops.split((torch.randn(ops.round_up(N, 64), K, generator=g) / math.sqrt(K)).to(cuda))
    bias = torch.randn(ops.round_up(N, 64), generator=g).to(cuda)
    res = ops.split(torch.randn(M, ops.round_up(N, 8), generator=g).to(cuda))
    outs = []
    for flag in (force, force | 1024):
        _lib.lib().hvr_debug_force_bn(flag)
        try:
            o, of, oT = ops.linear(a, w, N, bias=bias, relu=True, res=res, want_split=True, want_f32=True, want_T=True)
            torch.cuda.synchronize()
        finally:
            _lib.lib().hvr_debug_force_bn(0)
        outs.append((o.hi[:, :N].clone(), o.lo[:, :N].clone(), of[:, :N].clone(), oT.hi[:N, :M].clone(),
                     oT.lo[:N, :M].clone()))
    for x, y in zip(outs[0], outs[1]):
        assert torch.equal(x, y)
    ref = _ref_linear(a, w, N, bias, res, True, 1.0)
    assert _rel(outs[0][2].double().cpu(), ref) < 3e-5


@pytest.mark.parametrize('M,K,N,use_res', [
    (128 * 9 + 50, 256, 600, True),          # ragged M and N (600 = 9*64 + 24: clipped last chunk)
    (128 * 9 + 50, 256, 600, False),         # store-only staging
    (128 * 150 + 3, 128, 1024, True),        # more pair tiles than clusters: prefetch crosses tile boundaries
    (128 * 5, 1024, 320, True),              # odd number of 128-row tiles, N tail inside a 256 tile
])
def test_igemm_deep_epilogue_bit_identical(cuda, M, K, N, use_res):
    """The deep-epilogue pair kernel (2 ring stages, 3 in-place 64-column staging buffers per
    epilogue warp, residual prefetched two chunks ahead across tiles; flag 2048) writes the same
    bits as the 3-stage pair kernel (flag 4096) and as the per-row register epilogue (flag 1024)."""
    from hvrnet_b200 import _lib, ops
    g = torch.Generator().manual_seed(M + N + int(use_res))
    a = ops.split(torch.randn(M, K, generator=g).to(cuda))
    w = ops.split((torch.randn(ops.round_up(N, 64), K, generator=g) / math.sqrt(K)).to(cuda))
    bias = torch.randn(ops.round_up(N, 64), generator=g).to(cuda)
    res = ops.split(torch.randn(M, ops.round_up(N, 8), generator=g).to(cuda)) if use_res else None
    outs = []
    for flag in (512 | 2048, 512 | 4096, 512 | 4096 | 1024):
        _lib.lib().hvr_debug_force_bn(flag)
        try:
            o, _, _ = ops.linear(a, w, N, bias=bias, relu=True, res=res, want_split=True)
            torch.cuda.synchronize()
        finally:
            _lib.lib().hvr_debug_force_bn(0)
        outs.append((o.hi[:, :N].clone(), o.lo[:, :N].clone()))
    for other in outs[1:]:
        assert torch.equal(outs[0][0], other[0]) and torch.equal(outs[0][1], other[1])
    ref = _ref_linear(a, w, N, bias, res, True, 1.0)
    from hvrnet_b200.ops import Split
    assert _rel(ops.merge(Split(outs[0][0].contiguous(), outs[0][1].contiguous())).double().cpu(), ref) < 3e-5


@pytest.mark.parametrize('M,K,N,use_res,f32,flag', [
    (128 * 400 + 37, 256, 512, False, False, 0),       # lean pair kernel, 402 pair tiles for 74 clusters
    (128 * 400 + 37, 256, 512, True, False, 0),        # deep epilogue: the residual cursor reads the next tile ahead
    (128 * 400 + 37, 192, 520, True, True, 0),         # standard pair kernel (fp32 output), clipped N tile
    (128 * 700 + 5, 64, 128, False, False, 640),       # 256 x 128 pair tiles, one K step per tile (producer far ahead)
    (128 * 700 + 5, 64, 64, True, False, 512 | 2048),  # deep epilogue, ONE 64-column chunk per tile: cursor two tiles ahead
    (128 * 150, 128, 256, False, False, 0),            # exactly one wave + 1: 75 pair tiles
])
def test_igemm_cluster_launch_control_bit_identical(cuda, M, K, N, use_res, f32, flag):
    """Dynamic tile hand-out (bit 18 of the debug flags: grid = one cluster per pair tile, a running cluster cancels the
    launch of pending clusters and takes their tiles) against the static round-robin sequence: the same bits, every
    tile written exactly once (outputs are pre-filled with NaN)."""
    from hvrnet_b200 import _lib, ops
    from hvrnet_b200.ops import Split
    g = torch.Generator().manual_seed(M + N + K)
    a = ops.split(torch.randn(M, K, generator=g).to(cuda))
    w = ops.split((torch.randn(ops.round_up(N, 64), K, generator=g) / math.sqrt(K)).to(cuda))
    bias = torch.randn(ops.round_up(N, 64), generator=g).to(cuda)
    res = ops.split(torch.randn(M, ops.round_up(N, 8), generator=g).to(cuda)) if use_res else None
    outs = []
    for clc in (0, 1 << 18, 1 << 18):
        nan = torch.full((M, ops.round_up(N, 8)), float('nan'), device=cuda)
        out = Split(nan.bfloat16(), nan.bfloat16())
        _lib.lib().hvr_debug_force_bn(flag | clc)
        try:
            o, of, _ = ops.linear(a, w, N, bias=bias, relu=True, res=res, want_split=True, want_f32=f32, out=out)
            torch.cuda.synchronize()
        finally:
            _lib.lib().hvr_debug_force_bn(0)
        outs.append((o.hi[:, :N].clone(), o.lo[:, :N].clone(), of[:, :N].clone() if f32 else None))
    assert not bool(torch.isnan(outs[0][0].float()).any())
    for other in outs[1:]:
        assert torch.equal(outs[0][0].view(torch.int16), other[0].view(torch.int16))
        assert torch.equal(outs[0][1].view(torch.int16), other[1].view(torch.int16))
        if f32:
            assert torch.equal(outs[0][2], other[2])
    ref = _ref_linear(a, w, N, bias, res, True, 1.0)
    assert _rel(ops.merge(Split(outs[0][0].contiguous(), outs[0][1].contiguous())).double().cpu(), ref) < 3e-5


@pytest.mark.parametrize('tile', [(16, 8), (8, 16), (32, 4), (64, 2), (128, 1)])
def test_igemm_deep_epilogue_conv_tile_shapes(cuda, tile):
    """Deep epilogue on every pixel-box shape of the M tile: 1x1 conv + residual + ReLU on a map
    whose width / height are not multiples of the tile (clipped boxes)."""
    import torch.nn.functional as F
    from hvrnet_b200 import _lib, engine, ops
    from hvrnet_b200.ops import Split
    g = torch.Generator().manual_seed(100 + tile[0])
    B, H, W, C, N = 3, 38, 63, 64, 320
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(N, C, 1, 1, generator=g) / math.sqrt(C)
    r = torch.randn(B, N, H, W, generator=g)
    xs, rs = ops.nchw_to_nhwc_split(x.to(cuda)), ops.nchw_to_nhwc_split(r.to(cuda))
    wp = engine.pack_conv(w, None, cuda)
    rows = B * H * W
    outs = []
    for flag in (512 | 2048, 512 | 4096):
        out = Split.zeros((B, H, W, N), cuda)
        gd = ops.igemm_desc(xs, wp, N, taps=engine._taps(1, 1), out_whb=(W, H, B), tile=tile, relu=True,
                            res=Split(rs.hi.view(rows, N), rs.lo.view(rows, N)),
                            out=Split(out.hi.view(rows, N), out.lo.view(rows, N)))
        _lib.lib().hvr_debug_force_bn(flag)
        try:
            ops.igemm_run(gd)
            torch.cuda.synchronize()
        finally:
            _lib.lib().hvr_debug_force_bn(0)
        outs.append(out)
    assert torch.equal(outs[0].hi, outs[1].hi) and torch.equal(outs[0].lo, outs[1].lo)
    xm = ops.nhwc_split_to_nchw(xs).double().cpu()
    wm = ops.merge(wp).double().cpu()[:N].view(N, 1, 1, C).permute(0, 3, 1, 2)
    ref = (F.conv2d(xm, wm) + ops.nhwc_split_to_nchw(rs).double().cpu()).clamp_min(0)
    assert _rel(ops.nhwc_split_to_nchw(outs[0]).double().cpu(), ref) < 3e-5


@pytest.mark.parametrize('V,m,K,n,kind', [
    (7, 300, 1024, 4500, 'qk'),        # key-row queries: 3 M tiles per video (odd: idle half pair), ragged n
    (3, 448, 256, 600, 'qk'),          # even tiles per video
    (5, 300, 4544, 1024, 'pv'),        # P.V: B is a column block of X^T (batch stride = Npad columns)
    (2, 100, 192, 96, 'qk'),           # small: 1-CTA kernels
])
@pytest.mark.parametrize('force', [0, 64, 512])
def test_igemm_batched_b_matches_per_problem(cuda, V, m, K, n, kind, force):
    """ops.bmm (one launch, per-image B matrix through the third TMA coordinate; CTA pairs never
    straddle two images) == V separate ops.linear calls, bit for bit, and both match fp64."""
    from hvrnet_b200 import _lib, ops
    from hvrnet_b200.ops import Split
    g = torch.Generator().manual_seed(V * 100 + m + force)
    a = ops.split(torch.randn(V * m, K, generator=g).to(cuda))
    if kind == 'qk':                   # B_v = rows [v*npad, v*npad + n) of a [V*npad, K] matrix
        npad = ops.round_up(n, 64)
        b = ops.split((torch.randn(V * npad, K, generator=g) / math.sqrt(K)).to(cuda))
        stride = npad * b.hi.stride(0)
        bv = lambda v: Split(b.hi[v * npad:(v + 1) * npad], b.lo[v * npad:(v + 1) * npad])
    else:                              # B_v = columns [v*K, (v+1)*K) of a [n, V*K] matrix
        b = ops.split((torch.randn(n, V * K, generator=g) / math.sqrt(K)).to(cuda))
        stride = K
        bv = lambda v: Split(b.hi[:, v * K:(v + 1) * K], b.lo[:, v * K:(v + 1) * K])
    _lib.lib().hvr_debug_force_bn(force)
    try:
        o, of = ops.bmm(a, b, V, n, stride, alpha=0.25, want_split=True, want_f32=True)
        torch.cuda.synchronize()
        for v in range(V):
            av = a[v * m:(v + 1) * m]
            o1, f1, _ = ops.linear(av, bv(v), n, alpha=0.25, want_split=True, want_f32=True)
            torch.cuda.synchronize()
            assert torch.equal(of[v * m:(v + 1) * m, :n], f1[:, :n])
            assert torch.equal(o.hi[v * m:(v + 1) * m, :n], o1.hi[:, :n])
            assert torch.equal(o.lo[v * m:(v + 1) * m, :n], o1.lo[:, :n])
            ref = 0.25 * (ops.merge(Split(av.hi.contiguous(), av.lo.contiguous())).double().cpu()
                          @ ops.merge(Split(bv(v).hi.contiguous(), bv(v).lo.contiguous())).double().cpu()[:n].t())
            assert _rel(f1[:, :n].double().cpu(), ref) < 3e-5
        # the SIMT cross-check kernel evaluates the batched descriptor too
        _, of2 = ops.bmm(a, b, V, n, stride, alpha=0.25, want_split=False, want_f32=True, check_kernel=True)
        # (fp32 sequential accumulation over K up to 4544: looser than the tensor-core path itself)
        assert _rel(of2[:, :n].double().cpu(), of[:, :n].double().cpu()) < 1e-4
    finally:
        _lib.lib().hvr_debug_force_bn(0)


@pytest.mark.parametrize('B,H,W,C1,C2,N,stride,force', [
    (3, 38, 63, 512, 1024, 2048, 1, 0),      # layer4 block 0: conv3 (512) + downsample (1024)
    (2, 76, 126, 128, 256, 512, 2, 0),       # layer2 block 0: the downsample input is sampled with stride 2
    (1, 20, 24, 64, 64, 256, 1, 64),         # layer1 block 0 shapes on the 1-CTA kernel
    (2, 38, 63, 256, 72, 320, 1, 512),       # second operand with a K tail (72 = 64 + 8), ragged N, pair kernel
])
def test_igemm_second_a_operand(cuda, B, H, W, C1, C2, N, stride, force):
    """bn3(conv3(o)) + bn_d(conv_d(x)) as one contraction over K = [C1 | C2] (HvrIGemm.a2_*) ==
    the two-convolution evaluation in fp64 (resnet.py:243-255), and the SIMT cross-check kernel."""
    import torch.nn.functional as F
    from hvrnet_b200 import _lib, engine, ops
    g = torch.Generator().manual_seed(C1 + C2 + N)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    o = torch.randn(B, C1, Ho, Wo, generator=g)
    x = torch.randn(B, C2, H, W, generator=g)
    w3 = torch.randn(N, C1, 1, 1, generator=g) / math.sqrt(C1)
    wd = torch.randn(N, C2, 1, 1, generator=g) / math.sqrt(C2)
    bias = torch.randn(N, generator=g)
    os_, xs = ops.nchw_to_nhwc_split(o.to(cuda)), ops.nchw_to_nhwc_split(x.to(cuda))
    wcat = torch.cat([w3.reshape(N, C1), wd.reshape(N, C2)], 1).double()
    b = torch.zeros(ops.round_up(N, 64))
    b[:N] = bias
    cp = engine.ConvP(engine.pack_matrix(wcat, cuda), b.to(cuda), N, 1, C1, 1, cin2=C2)
    _lib.lib().hvr_debug_force_bn(force)
    try:
        out, _ = engine.conv(os_, cp, relu=True, a2=xs, a2_stride=stride)
        chk, _ = engine.conv(os_, cp, relu=True, a2=xs, a2_stride=stride, check_kernel=True)
        torch.cuda.synchronize()
    finally:
        _lib.lib().hvr_debug_force_bn(0)
    om, xm = ops.nhwc_split_to_nchw(os_).double().cpu(), ops.nhwc_split_to_nchw(xs).double().cpu()
    wm = ops.merge(cp.w).double().cpu()[:N]
    ref = F.conv2d(om, wm[:, :C1].reshape(N, C1, 1, 1)) + \
        F.conv2d(xm, wm[:, C1:C1 + C2].reshape(N, C2, 1, 1), stride=stride) + bias.double().view(1, -1, 1, 1)
    ref = ref.clamp_min(0)
    assert _rel(ops.nhwc_split_to_nchw(out)[:, :N].double().cpu(), ref) < 3e-5
    assert _rel(ops.nhwc_split_to_nchw(chk)[:, :N].double().cpu(), ref) < 3e-5


@pytest.mark.parametrize('tile', [(16, 8), (8, 16), (32, 4), (64, 2), (128, 1)])
def test_igemm_conv_tile_shapes(cuda, tile):
    """Every pixel-box shape of the M tile (and of the TMA epilogue box) on a 3x3 conv with residual."""
    import torch.nn.functional as F
    from hvrnet_b200 import engine, ops
    from hvrnet_b200.ops import Split
    g = torch.Generator().manual_seed(tile[0])
    B, H, W, C, N = 2, 21, 37, 64, 96
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(N, C, 3, 3, generator=g) / math.sqrt(C * 9)
    r = torch.randn(B, N, H, W, generator=g)
    xs, rs = ops.nchw_to_nhwc_split(x.to(cuda)), ops.nchw_to_nhwc_split(r.to(cuda))
    wp = engine.pack_conv(w, None, cuda)
    out = Split.empty((B, H, W, N), cuda)
    rows = B * H * W
    gd = ops.igemm_desc(xs, wp, N, taps=engine._taps(3, 1), out_whb=(W, H, B), tile=tile, relu=True,
                        res=Split(rs.hi.view(rows, N), rs.lo.view(rows, N)),
                        out=Split(out.hi.view(rows, N), out.lo.view(rows, N)))
    ops.igemm_run(gd)
    torch.cuda.synchronize()
    xm = ops.nhwc_split_to_nchw(xs).double().cpu()
    wm = ops.merge(wp).double().cpu()[:N].view(N, 3, 3, C).permute(0, 3, 1, 2)
    ref = (F.conv2d(xm, wm, padding=1) + ops.nhwc_split_to_nchw(rs).double().cpu()).clamp_min(0)
    assert _rel(ops.nhwc_split_to_nchw(out).double().cpu(), ref) < 3e-5


def Split_rows(s):
    from hvrnet_b200.ops import Split
    return Split(s.hi.contiguous(), s.lo.contiguous())


def test_igemm_k_tail_and_single_pass(cuda):
    """K = 200 is not a multiple of the 64-wide K tile: TMA zero-fills the tail (the P.V
    contraction has K = number of proposals)."""
    from hvrnet_b200 import ops
    g = torch.Generator().manual_seed(5)
    M, K, N = 150, 200, 96
    A = torch.zeros(M, 256)
    A[:, :K] = torch.randn(M, K, generator=g)
    B = torch.zeros(128, 256)
    B[:N, :K] = torch.randn(N, K, generator=g)
    a, b = ops.split(A.to(cuda)), ops.split(B.to(cuda))
    from hvrnet_b200.ops import Split
    av = Split(a.hi[:, :K], a.lo[:, :K])
    bv = Split(b.hi[:, :K], b.lo[:, :K])
    _, of, _ = ops.linear(av, bv, N, want_split=False, want_f32=True)
    ref = (A[:, :K].double() @ B[:N, :K].double().t())
    assert _rel(of[:, :N].double().cpu(), ref) < 3e-5
    _, o1, _ = ops.linear(av, bv, N, want_split=False, want_f32=True, passes=1)
    refh = a.hi[:, :K].double().cpu() @ b.hi[:N, :K].double().cpu().t()
    assert _rel(o1[:, :N].double().cpu(), refh) < 1e-5      # single product: exact up to fp32 accumulation


@pytest.mark.parametrize('B,H,W,C,N,k,dil,stride', [
    (2, 38, 63, 128, 64, 3, 2, 1),      # layer4-style dilated 3x3
    (1, 19, 31, 64, 128, 3, 1, 1),      # 3x3 pad 1
    (2, 38, 63, 256, 128, 1, 1, 2),     # strided 1x1 (caffe bottleneck conv1 / downsample)
    (1, 20, 24, 64, 64, 1, 1, 1),
])
def test_igemm_conv(cuda, B, H, W, C, N, k, dil, stride):
    import torch.nn.functional as F
    from hvrnet_b200 import engine, ops
    g = torch.Generator().manual_seed(B * H + C + N + k)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(N, C, k, k, generator=g) / math.sqrt(C * k * k)
    bias = torch.randn(N, generator=g)
    xs = ops.nchw_to_nhwc_split(x.to(cuda))
    cp = engine.ConvP(engine.pack_conv(w, None, cuda), bias.to(cuda), N, k, C, dil)
    res = torch.randn(B, N, (H - 1) // stride + 1, (W - 1) // stride + 1, generator=g)
    rs = ops.nchw_to_nhwc_split(res.to(cuda))
    out, of = engine.conv(xs, cp, stride=stride, relu=True, res=rs, want_f32=True)
    torch.cuda.synchronize()
    xm = ops.nhwc_split_to_nchw(xs).double().cpu()
    wm = ops.merge(cp.w).double().cpu()[:N].view(N, k, k, C).permute(0, 3, 1, 2)
    ref = F.conv2d(xm, wm, bias.double(), stride=stride, padding=dil * (k // 2), dilation=dil)
    ref = (ref + ops.nhwc_split_to_nchw(rs).double().cpu()).clamp_min(0)
    got = of[..., :N].permute(0, 3, 1, 2).double().cpu()
    assert float((got - ref).abs().max() / ref.abs().max()) < 3e-5
    got2 = ops.nhwc_split_to_nchw(out).double().cpu()
    assert float((got2 - ref).abs().max() / ref.abs().max()) < 3e-5


def test_stem_and_maxpool(cuda):
    import torch.nn.functional as F
    from hvrnet_b200 import engine, ops
    g = torch.Generator().manual_seed(3)
    img = torch.randn(1, 3, 70, 90, generator=g) * 50
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.02
    col = ops.im2col_stem(img.to(cuda))
    wp = engine.pack_matrix(w.permute(0, 2, 3, 1).reshape(64, 147), cuda, 64, 192)
    cp = engine.ConvP(wp, None, 64, 1, 192)
    x, _ = engine.conv(col, cp, relu=True)
    y = ops.maxpool3x3s2(x)
    ref = F.max_pool2d(F.relu(F.conv2d(img.double(), w.double(), stride=2, padding=3)), 3, 2, 1)
    got = ops.nhwc_split_to_nchw(y).double().cpu()
    assert got.shape == ref.shape
    assert float((got - ref).abs().max() / ref.abs().max()) < 5e-5


# ---------------------------------------------------------------------------------------
# RoIAlign: bit-exact against the C oracle
# ---------------------------------------------------------------------------------------
def _rois(g, n, n_imgs, w=1000., h=600.):
    x1 = torch.rand(n, generator=g) * w
    y1 = torch.rand(n, generator=g) * h
    bw = torch.rand(n, generator=g) ** 2 * w * 0.6
    bh = torch.rand(n, generator=g) ** 2 * h * 0.6
    r = torch.stack([torch.randint(0, n_imgs, (n,), generator=g).float(), x1, y1,
                     (x1 + bw).clamp(max=w - 1), (y1 + bh).clamp(max=h - 1)], 1)
    r[0, 1:] = torch.tensor([-40., -30., 10., 12.])          # partly outside: OOB samples
    r[1, 1:] = torch.tensor([990., 590., 1050., 640.])       # beyond the map: all-zero bins
    r[2, 1:] = torch.tensor([100., 100., 100., 100.])        # degenerate 1-pixel roi
    r[3, 1:] = torch.tensor([300., 200., 200., 100.])        # x2 < x1: width clamps to 0
    return r


def test_roi_align_bit_exact(cuda):
    from hvrnet_b200 import ops
    from oracle import cref
    g = torch.Generator().manual_seed(11)
    feat = torch.randn(3, 256, 38, 63, generator=g)
    rois = _rois(g, 120, 3)
    ref = cref.roi_align(feat, rois)
    out = ops.roi_align(feat.to(cuda), rois.to(cuda))
    assert torch.equal(out.cpu(), ref)
    # NHWC in / NHWC out + split copy (the pipeline's variant)
    fn = ops.nchw_to_nhwc(feat.to(cuda))
    o2, sp = ops.roi_align(fn, rois.to(cuda), feat_nhwc=True, out_nhwc=True, want_split=True)
    assert torch.equal(o2.permute(0, 3, 1, 2).cpu(), ref)
    m = ops.merge(sp).view(-1, 7, 7, 256).permute(0, 3, 1, 2).cpu()
    assert float((m - ref).abs().max()) <= float(ref.abs().max()) * 2.0 ** -16


def test_roi_align_vs_reference_cuda_op(cuda):
    """R8 against the REFERENCE's own CUDA op: mmdet/ops/roi_align/src/roi_align_cuda.cpp + roi_align_kernel.cu,
    compiled unmodified for sm_100a in the build container (oracle/build.py::build_ref_roi_align, output in
    oracle/_ref, shipped to the GPU box) and called through its pybind `forward` exactly as roi_align.py:12-30
    does.  Built with -fmad=false (every product and sum rounded - the contract of this repo's kernels and of
    oracle/c) the reference kernel, both of our kernels and the C oracle agree BIT FOR BIT; against the default
    build (nvcc contracts a*b+c into FMAs) the results agree to 1e-5 relative."""
    from hvrnet_b200 import ops
    from oracle import build, cref
    ref_fma, ref_strict = build.load_ref_roi_align(True), build.load_ref_roi_align(False)
    if ref_fma is None or ref_strict is None:
        pytest.skip('oracle/_ref reference RoIAlign op not built (needs /root/reference at build time)')
    g = torch.Generator().manual_seed(17)
    feat = torch.randn(3, 256, 38, 63, generator=g)
    rois = _rois(g, 400, 3)
    f, r = feat.to(cuda), rois.to(cuda)

    def reference(mod, out_size, scale, sn, ff=f, rr=r):
        out = ff.new_zeros(rr.shape[0], ff.shape[1], out_size, out_size)        # roi_align.py:23
        assert mod.forward(ff, rr, out_size, out_size, scale, sn, out) == 1
        torch.cuda.synchronize()
        return out
    strict = reference(ref_strict, 7, 1 / 16., 2)
    ours = ops.roi_align(f, r)                                                  # reference layout (NCHW)
    assert torch.equal(ours.view(torch.int32), strict.view(torch.int32))
    rows = ops.roi_align(ops.nchw_to_nhwc(f), r, feat_nhwc=True, out_nhwc=True)  # the pipeline's sn2 kernel
    assert torch.equal(rows.permute(0, 3, 1, 2).contiguous().view(torch.int32), strict.view(torch.int32))
    assert torch.equal(cref.roi_align(feat, rois).view(torch.int32), strict.cpu().view(torch.int32))
    assert _rel(ours, reference(ref_fma, 7, 1 / 16., 2)) < 1e-5
    # adaptive sampling (sample_num = 0) and another geometry: the generic kernel
    f2 = torch.randn(2, 16, 15, 15, generator=g).to(cuda)
    r2 = torch.tensor([[0, 0, 0, 50, 50], [0, 10, 30, 43, 55], [1, 67, 40, 110, 120]], dtype=torch.float32).to(cuda)
    for sn in (0, 2, 3):
        a = ops.roi_align(f2, r2, out_size=3, spatial_scale=1 / 8., sample_num=sn)
        b = reference(ref_strict, 3, 1 / 8., sn, f2, r2)
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), sn


@pytest.mark.parametrize('C,out_size', [(256, 7), (128, 7), (512, 7), (256, 3), (16, 7), (40, 3)])
def test_roi_align_sn2_kernel_equals_generic_per_bin_kernel(cuda, C, out_size):
    """roi_align_sn2_kernel (taps shared by the two y-samples of a bin reused from registers; RoIs with
    samples outside the map on its per-sample path) against roi_align_kernel<false> (16 loads per output
    vector) and the oracle: identical bits, fp32 rows and split rows.  C = 16: a warp spans several bins;
    C = 40 has no sn2 launch shape (256 % (C/4) != 0) and checks the fall-back."""
    from hvrnet_b200 import _lib, ops
    from oracle import cref
    g = torch.Generator().manual_seed(21 + C)
    feat = torch.randn(2, 38, 63, C, generator=g)
    rois = _rois(g, 300, 2)
    rois[4:40, 3:] = rois[4:40, 1:3] + torch.rand(36, 2, generator=g) * 60      # small RoIs: heavy reuse
    rois[40, 1:] = torch.tensor([0., 0., 999., 599.])                            # whole frame: no reuse
    rois[41, 1:] = torch.tensor([-17., 300., 5., 320.])
    ref = cref.roi_align(feat, rois, out_size=out_size, feat_nhwc=True, out_nhwc=True)
    f, r = feat.to(cuda), rois.to(cuda)
    outs = []
    try:
        for variant in (0, 1):
            assert _lib.lib().hvr_debug_roi_variant(variant) == 0
            o, sp = ops.roi_align(f, r, out_size=out_size, feat_nhwc=True, out_nhwc=True, want_split=True)
            outs.append((o.cpu(), sp.hi.cpu(), sp.lo.cpu()))
    finally:
        _lib.lib().hvr_debug_roi_variant(0)
    assert torch.equal(outs[0][0].view(torch.int32), ref.view(torch.int32))
    bits = lambda t: t.view(torch.int16 if t.dtype == torch.bfloat16 else torch.int32)
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(bits(a), bits(b))


@pytest.mark.parametrize('C,out_size', [(256, 7), (128, 7), (256, 3), (16, 7), (40, 3), (512, 7)])
def test_roi_align_fast_variant_vs_strict(cuda, C, out_size):
    """hvr_roi_align_fwd_fast (roi_align_sep_kernel / roi_align_slab_kernel: separable, fused multiply-add) against
    the strict kernel and the C oracle.  Tolerance (the contract in include/hvr_b200.h): every element within 1e-5
    of the largest |reference| value of its RoI; bins the oracle evaluates to exactly 0 are exactly 0; the split
    rows are the split of the fp32 rows; both launch shapes give the same bits."""
    from hvrnet_b200 import ops
    from oracle import cref
    g = torch.Generator().manual_seed(31 + C)
    feat = torch.randn(2, 38, 63, C, generator=g)
    rois = _rois(g, 300, 2)
    rois[4:40, 3:] = rois[4:40, 1:3] + torch.rand(36, 2, generator=g) * 60
    rois[40, 1:] = torch.tensor([0., 0., 999., 599.])
    rois[41, 1:] = torch.tensor([-17., 300., 5., 320.])
    rois[42, 1:] = 0.                                                            # the pad RoI of a batched window
    rois[43, 1:] = torch.tensor([1500., 100., 1600., 200.])                      # entirely outside the map: all zero
    rois[44, 1:] = torch.tensor([100., -500., 200., -100.])
    rois[45, 1:] = torch.tensor([48., 100., 495., 300.])        # x samples exactly on columns 4, 6, 8 ...: 2 taps with a hole
    rois[46, 1:] = torch.tensor([44., 50., 603., 400.])         # x samples at 4.0 and 6.5: 3 taps with a hole
    ref = cref.roi_align(feat, rois, out_size=out_size, feat_nhwc=True, out_nhwc=True)
    f, r = feat.to(cuda), rois.to(cuda)
    from hvrnet_b200 import _lib
    assert _lib.lib().hvr_debug_roi_variant(6) == 0                               # the RoI-per-CTA launch shape,
    _lib.lib().hvr_debug_roi_variant(8)                                           # 4 channels per thread
    try:
        o, sp = ops.roi_align(f, r, out_size=out_size, feat_nhwc=True, out_nhwc=True, want_split=True, arithmetic='fast')
        # 8 channels per thread: the same operations per element -> the same bits
        _lib.lib().hvr_debug_roi_variant(7)
        o_8, sp_8 = ops.roi_align(f, r, out_size=out_size, feat_nhwc=True, out_nhwc=True, want_split=True, arithmetic='fast')
        assert torch.equal(o_8.view(torch.int32), o.view(torch.int32)) and torch.equal(sp_8.lo.view(torch.int16), sp.lo.view(torch.int16))
        # the 8-channel kernel's variants at C = 256: 10 adjacent channels, 11 lane-interleaved, 12 = 16 per thread (all with the
        # two-row-cache walk); 13 = lane-interleaved with the per-RoI row program (default, = o_8 above), 14 = 13 + L1 prefetch, 15 = 13 built for 3 CTAs per SM
        for layout in (10, 11, 12, 14, 15):
            _lib.lib().hvr_debug_roi_variant(layout)
            o_l, sp_l = ops.roi_align(f, r, out_size=out_size, feat_nhwc=True, out_nhwc=True, want_split=True, arithmetic='fast')
            assert torch.equal(o_l.view(torch.int32), o.view(torch.int32)) and torch.equal(sp_l.hi.view(torch.int16), sp.hi.view(torch.int16))
        _lib.lib().hvr_debug_roi_variant(13)
        _lib.lib().hvr_debug_roi_variant(8)
        # the slab launch shape (CTA = frame x 16-channel slab, map slab in shared memory): the same separable core,
        # so identical bits wherever it applies (C % 16 == 0, out_size <= 8)
        _lib.lib().hvr_debug_roi_variant(5)
        o_s, sp_s = ops.roi_align(f, r, out_size=out_size, feat_nhwc=True, out_nhwc=True, want_split=True, arithmetic='fast')
    finally:
        _lib.lib().hvr_debug_roi_variant(7)
        _lib.lib().hvr_debug_roi_variant(13)
        _lib.lib().hvr_debug_roi_variant(4)
    if C == 512:
        # no RoI-per-CTA launch shape (7 * 128 threads > 448): with the slab shape disabled this is the strict kernel;
        # the slab shape still applies and is held to the tolerance below
        assert torch.equal(o.cpu().view(torch.int32), ref.view(torch.int32))
        o, sp = o_s, sp_s
    elif C % 16 == 0:
        assert torch.equal(o_s.view(torch.int32), o.view(torch.int32))
        assert torch.equal(sp_s.hi.view(torch.int16), sp.hi.view(torch.int16)) and torch.equal(sp_s.lo.view(torch.int16), sp.lo.view(torch.int16))
    o = o.cpu()
    scale = ref.abs().flatten(1).amax(1).clamp_min(1e-30).view(-1, 1, 1, 1)
    assert float(((o.double() - ref.double()).abs() / scale).max()) < 1e-5
    dead = ref.abs().flatten(1).amax(1) == 0
    assert bool(dead.any()) and not bool(o[dead].any())
    sp2 = ops.split(o.to(cuda).view(o.shape[0], -1))
    assert torch.equal(sp.hi.view(torch.int16), sp2.hi.view(torch.int16))
    assert torch.equal(sp.lo.view(torch.int16), sp2.lo.view(torch.int16))
    # split rows only (the pipeline's call) give the same bits
    _lib.lib().hvr_debug_roi_variant(5 if C == 512 else 6)               # (default channel grouping: 8 per thread)
    try:
        _, sp3 = ops.roi_align(f, r, out_size=out_size, feat_nhwc=True, out_nhwc=True, want_split=True, want_f32=False,
                               arithmetic='fast')
    finally:
        _lib.lib().hvr_debug_roi_variant(4)
    assert torch.equal(sp.hi.view(torch.int16), sp3.hi.view(torch.int16))


def test_roi_align_fast_variant_full_size(cuda):
    """BASELINE.json size (15 frames x 300 proposals, 38x63x256 maps, one launch) through the fast variant:
    1e-5 of each RoI's largest value against the C oracle on a sample of the RoIs."""
    from hvrnet_b200 import ops
    from oracle import cref
    g = torch.Generator().manual_seed(13)
    T, P = 15, 300
    feat = torch.randn(T, 38, 63, 256, generator=g)
    rois = _rois(g, T * P, T)
    rois[:, 0] = torch.arange(T * P) // P
    o = ops.roi_align(feat.to(cuda), rois.to(cuda), feat_nhwc=True, out_nhwc=True, arithmetic='fast')
    pick = torch.arange(0, T * P, 37)
    ref = cref.roi_align(feat, rois[pick], feat_nhwc=True, out_nhwc=True)
    scale = ref.abs().flatten(1).amax(1).clamp_min(1e-30).view(-1, 1, 1, 1)
    assert float(((o[pick].cpu().double() - ref.double()).abs() / scale).max()) < 1e-5


def test_roi_align_full_size(cuda):
    """BASELINE.json size: 15 frames x 300 proposals on 38x63x256 maps in ONE launch, checked
    bit-exactly against the C oracle on a sample of the RoIs."""
    from hvrnet_b200 import ops
    from oracle import cref
    g = torch.Generator().manual_seed(12)
    T, P = 15, 300
    feat = torch.randn(T, 38, 63, 256, generator=g)
    rois = _rois(g, T * P, T)
    rois[:, 0] = torch.arange(T * P) // P
    o = ops.roi_align(feat.to(cuda), rois.to(cuda), feat_nhwc=True, out_nhwc=True)
    pick = torch.arange(0, T * P, 37)
    ref = cref.roi_align(feat, rois[pick], feat_nhwc=True, out_nhwc=True)
    assert torch.equal(o[pick].cpu(), ref)


def test_roi_align_gradcheck_recipe_and_empty(cuda):
    """Input recipe of mmdet/ops/roi_align/gradcheck.py:11-30 (15x15 maps, scale 1/8)."""
    from hvrnet_b200 import ops
    from oracle import cref
    g = torch.Generator().manual_seed(2)
    feat = torch.randn(2, 16, 15, 15, generator=g)
    rois = torch.tensor([[0, 0, 0, 50, 50], [0, 10, 30, 43, 55], [1, 67, 40, 110, 120]], dtype=torch.float32)
    ref = cref.roi_align(feat, rois, out_size=2, spatial_scale=1 / 8., sample_num=2)
    out = ops.roi_align(feat.to(cuda), rois.to(cuda), out_size=2, spatial_scale=1 / 8., sample_num=2)
    assert torch.equal(out.cpu(), ref)
    ref0 = cref.roi_align(feat, rois, out_size=3, spatial_scale=1 / 8., sample_num=0)     # adaptive sampling
    out0 = ops.roi_align(feat.to(cuda), rois.to(cuda), out_size=3, spatial_scale=1 / 8., sample_num=0)
    assert torch.equal(out0.cpu(), ref0)
    e = ops.roi_align(feat.to(cuda), rois[:0].to(cuda))
    assert e.shape == (0, 16, 7, 7)


# ---------------------------------------------------------------------------------------
# NMS: bit-exact indices
# ---------------------------------------------------------------------------------------
def _dets(g, n, spread=400.):
    c = torch.rand(n, 2, generator=g) * spread
    wh = torch.rand(n, 2, generator=g) * 120 + 4
    s = torch.rand(n, generator=g)
    return torch.cat([c - wh / 2, c + wh / 2, s[:, None]], 1)


def test_nms_doctest(cuda):
    """mmdet/ops/nms/nms_wrapper.py:25-35."""
    from hvrnet_b200 import ops
    d = torch.tensor([[49.1, 32.4, 51.0, 35.9, 0.9], [49.3, 32.9, 51.0, 35.3, 0.9], [49.2, 31.8, 51.0, 35.4, 0.5],
                      [35.1, 11.5, 39.1, 15.7, 0.5], [35.6, 11.8, 39.3, 14.2, 0.5], [35.3, 11.5, 39.9, 14.5, 0.4],
                      [35.2, 11.7, 39.7, 15.7, 0.3]])
    assert ops.nms(d.to(cuda), 0.7).cpu().tolist() == [0, 3, 4]
    assert ops.nms(d[:0].to(cuda), 0.7).numel() == 0


@pytest.mark.parametrize('n,thr', [(1, 0.5), (63, 0.3), (64, 0.5), (65, 0.7), (1000, 0.3), (6000, 0.7)])
def test_nms_bit_exact(cuda, n, thr):
    from hvrnet_b200 import ops
    from oracle import cref
    g = torch.Generator().manual_seed(n)
    d = _dets(g, n, spread=300. if n < 2000 else 900.)
    d[::7, 4] = d[0, 4]                                   # score ties: total order = index ascending
    for strict in (True, False):
        ref = cref.nms(d, thr, strict_gt=strict)
        got = ops.nms(d.to(cuda), thr, strict_gt=strict).cpu()
        assert torch.equal(got, ref)


@pytest.mark.parametrize('n,thr', [(1, 0.5), (64, 0.5), (65, 0.7), (1000, 0.3), (6000, 0.7)])
def test_nms_vs_reference_cuda_op(cuda, n, thr):
    """R7 against the REFERENCE's own CUDA op: mmdet/ops/nms/src/nms_cuda.cpp + nms_kernel.cu compiled unmodified
    for sm_100a in the build container (oracle/build.py::build_ref_nms_cuda -> oracle/_ref) and called through its
    pybind `nms` as nms_wrapper.py:50-58 does: same kept indices (strict `>`, ascending), without score ties
    (the reference's device sort is not stable)."""
    from hvrnet_b200 import ops
    from oracle import build, cref
    ref = build.load_ref_nms_cuda()
    if ref is None:
        pytest.skip('oracle/_ref reference NMS CUDA op not built (needs /root/reference at build time)')
    g = torch.Generator().manual_seed(100 + n)
    d = _dets(g, n, spread=300. if n < 2000 else 900.)
    d[:, 4] = torch.randperm(n, generator=g).float() / n            # distinct scores
    keep_ref = ref.nms(d.to(cuda), thr)
    torch.cuda.synchronize()
    assert keep_ref.dtype == torch.int64
    assert torch.equal(ops.nms(d.to(cuda), thr).cpu(), keep_ref.cpu())
    assert torch.equal(cref.nms(d, thr, strict_gt=True), keep_ref.cpu())


# ---------------------------------------------------------------------------------------
# RPN proposals: bit-exact anchor indices, boxes to 1e-4 px
# ---------------------------------------------------------------------------------------
def test_rpn_proposals(cuda):
    from hvrnet_b200 import ops
    from oracle import ref_torch as R
    g = torch.Generator().manual_seed(21)
    T, H, W, A = 3, 38, 63, 12
    cls = torch.randn(T, A, H, W, generator=g) * 3
    reg = torch.randn(T, 4 * A, H, W, generator=g) * 0.3
    cls[0, :, :4, :4] = 40.0                              # saturated sigmoid: ties in score, not in logit order
    base = R.gen_base_anchors(16, (4, 8, 16, 32), (0.5, 1.0, 2.0))
    anchors = R.grid_anchors(base, (H, W), 16)
    packed = torch.cat([cls.permute(0, 2, 3, 1), reg.permute(0, 2, 3, 1),
                        torch.zeros(T, H, W, 4)], -1).contiguous().to(cuda)      # [T,H,W,64] like the RPN GEMM output
    props, counts, idx = ops.rpn_proposals(packed, packed.view(-1)[A:], 64, 64, T, H, W, A, base.to(cuda), 16,
                                           (600, 1000), want_idx=True)
    for t in range(T):
        ref, ridx = R.rpn_proposals_single(cls[t], reg[t], anchors, (600, 1000), return_aux=True)
        k = int(counts[t])
        assert k == ref.shape[0]
        assert torch.equal(idx[t, :k].cpu().long(), ridx)
        assert float((props[t, :k].cpu() - ref).abs().max()) < 1e-3
        assert float((props[t, :k, 4].cpu() - ref[:, 4]).abs().max()) < 1e-6
        assert float(props[t, k:].abs().sum()) == 0


# ---------------------------------------------------------------------------------------
# detection post-processing
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize('n,scale,rescale,boost', [(300, 1.0, False, 0.0), (300, 1.6, True, 0.0), (37, 1.0, True, 0.0),
                                                   (300, 1.0, False, 6.0)])
def test_det_postprocess(cuda, n, scale, rescale, boost):
    from hvrnet_b200 import ops
    from oracle import ref_torch as R
    g = torch.Generator().manual_seed(n + int(scale * 10))
    d = _dets(g, n, spread=500.)
    rois = torch.cat([torch.zeros(n, 1), d[:, :4]], 1)
    cls = torch.randn(n, 31, generator=g) * 2
    cls[:, 1:] += boost - 3.0            # boost>0: many classes pass the threshold -> >300 candidates -> top-k branch
    reg = torch.randn(n, 4, generator=g) * 0.5
    dets_ref, labels_ref = R.get_det_bboxes(rois, [cls], [reg], (600, 1000), scale, rescale)
    dets, labels, nd = ops.det_postprocess(rois.to(cuda), cls.to(cuda), reg.to(cuda), (600, 1000), scale, rescale)
    k = int(nd)
    assert k == dets_ref[0].shape[0]
    assert torch.equal(labels[:k].cpu(), labels_ref[0])
    assert float((dets[:k, :4].cpu() - dets_ref[0][:, :4]).abs().max()) < 1e-3
    assert float((dets[:k, 4].cpu() - dets_ref[0][:, 4]).abs().max()) < 1e-6


@pytest.mark.parametrize('G,n,rescale', [(7, 300, True), (3, 37, False), (1, 300, True)])
def test_det_postprocess_batched_bit_identical(cuda, G, n, rescale):
    """G problems through one launch per stage (64-bit composite sort keys, one scan over all
    problems) == G calls of the per-problem path, bit for bit; problems alternate between the
    concatenation branch (<= 300 candidates) and the top-k branch (> 300), and the inputs are
    strided views of one [G*n, 64] head output as the runtime passes them."""
    from hvrnet_b200 import ops
    g = torch.Generator().manual_seed(G * 1000 + n)
    d = _dets(g, G * n, spread=500.)
    rois = torch.cat([torch.zeros(G * n, 1), d[:, :4]], 1).to(cuda)
    out = torch.randn(G * n, 64, generator=g) * 2
    for p_ in range(G):
        out[p_ * n:(p_ + 1) * n, 1:31] += 3.0 if p_ % 2 else -14.0    # odd: > 300 candidates, even: a few dozen
    out[:, 31:35] *= 0.25
    out = out.to(cuda)
    cls, reg = out[:, :31], out[:, 31:35]
    D, L, K = ops.det_postprocess_batched(rois, cls, reg, G, (600, 1000), 1.6, rescale, n_cls=31)
    branches = set()
    for p_ in range(G):
        sl = slice(p_ * n, (p_ + 1) * n)
        d1, l1, k1 = ops.det_postprocess(rois[sl], cls[sl], reg[sl], (600, 1000), 1.6, rescale, n_cls=31)
        k = int(k1)
        assert int(K[p_]) == k
        assert torch.equal(D[p_, :k], d1[:k]) and torch.equal(L[p_, :k], l1[:k])
        branches.add(k == 300)       # 300 = max_per_img reached (top-k branch)
    if G > 1 and n == 300:
        assert branches == {True, False}


@pytest.mark.parametrize('M,D', [(4544, 1024), (300, 1024), (129, 72), (64, 8)])
def test_transpose_split(cuda, M, D):
    """X^T operand of the P.V products: exact transposed copy, pad columns [M, round_up(M,64)) zero."""
    from hvrnet_b200 import ops
    g = torch.Generator().manual_seed(M + D)
    s = ops.split(torch.randn(M, D, generator=g).to(cuda))
    t = ops.transpose_split(s)
    ld = ops.round_up(M, 64)
    assert t.shape == (D, ld)
    assert torch.equal(t.hi[:, :M], s.hi.t()) and torch.equal(t.lo[:, :M], s.lo.t())
    assert int(t.hi[:, M:].view(torch.int16).abs().sum()) == 0 and int(t.lo[:, M:].view(torch.int16).abs().sum()) == 0
    # a row-strided input view (the first n columns of a wider matrix)
    w = ops.split(torch.randn(M, D + 64, generator=g).to(cuda))
    t2 = ops.transpose_split(w, D)
    assert torch.equal(t2.hi[:, :M], w.hi[:, :D].t()) and torch.equal(t2.lo[:, :M], w.lo[:, :D].t())


def test_softmax_rows(cuda):
    from hvrnet_b200 import ops
    g = torch.Generator().manual_seed(9)
    for rows, cols in [(5, 17), (300, 4500), (7, 9000)]:
        S = torch.randn(rows, ops.round_up(cols, 4), generator=g) * 3
        P = ops.softmax_rows_split(S.to(cuda), cols)
        ref = torch.softmax(S[:, :cols].double(), 1)
        got = ops.merge(P)[:, :cols].double().cpu()
        assert float((got - ref).abs().max() / ref.max()) < 1e-5
        assert float(ops.merge(P)[:, cols:].abs().sum()) == 0


@pytest.mark.parametrize('h,w', [(720, 1280), (480, 640), (240, 320), (600, 1000), (333, 777)])
def test_preprocess_bit_exact(cuda, h, w):
    """Next row N2: fused resize + normalise + pad + CHW vs the numpy oracle (itself pinned to cv2)."""
    import numpy as np
    from hvrnet_b200 import preprocess as pp
    from oracle import preprocess as P
    rng = np.random.default_rng(h * 7 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref, rmeta = P.preprocess(img)
    out, meta = pp.preprocess(torch.from_numpy(img).to(cuda))
    assert meta == rmeta
    assert out.shape == (1,) + ref.shape
    assert np.array_equal(out[0].cpu().numpy(), ref)


# ---------------------------------------------------------------------------------------
# next row N4: video descriptors and similarity-based support selection
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize('V,T,hw,C', [(7, 15, (38, 63), 256), (3, 2, (5, 7), 40), (1, 1, (1, 1), 8)])
def test_video_descriptor(cuda, V, T, hw, C):
    """hvr_video_descriptor against adaptive_avg_pool2d + max over the frames (hnmb_rcnn.py:78-81):
    fp32 sums in a different order, so 1e-5 relative; run twice -> identical bits (fixed order)."""
    from hvrnet_b200 import ops
    from oracle import ref_torch as R
    g = torch.Generator().manual_seed(31)
    c5 = (torch.randn(V * T, C, hw[0], hw[1], generator=g) + torch.rand(V * T, C, 1, 1, generator=g) * 2).clamp(min=0)
    ref = torch.stack([R.video_descriptor(c5[v * T:(v + 1) * T]) for v in range(V)])
    x = c5.permute(0, 2, 3, 1).contiguous().to(cuda)
    d = ops.video_descriptor(x, V)
    assert d.shape == (V, C) and _rel(d.cpu(), ref) < 1e-5
    assert torch.equal(d, ops.video_descriptor(x, V))


@pytest.mark.parametrize('G,C,n', [(300, 256, 4), (5, 64, 4), (2, 256, 1), (1, 16, 4), (700, 32, 8)])
def test_support_select(cuda, G, C, n):
    """hvr_support_select against the oracle's select_support_by_similarity (softmax of scaled dot products
    over the other videos, hnmb_rcnn.py:85-88; largest first, ties to the lower index): indices exactly,
    weights to 1e-5; -1 where fewer than n other videos exist."""
    from hvrnet_b200 import ops
    from oracle import ref_torch as R
    g = torch.Generator().manual_seed(41 + G)
    desc = torch.randn(G, C, generator=g).abs() * (8.0 / C ** 0.5)
    g0, nl = (G // 3, min(G - G // 3, 9))
    idx, w = ops.support_select(desc.to(cuda), g0, nl, n, want_weights=True)
    for i in range(nl):
        ri, rw = R.select_support_by_similarity(desc, g0 + i, n)
        assert idx[i].cpu().tolist() == ri + [-1] * (n - len(ri))
        got = torch.cat([w[i, :g0 + i], w[i, g0 + i + 1:]]).cpu()
        assert got.numel() == 0 or float((got - rw).abs().max()) < 1e-5
    # exact ties: identical descriptors -> the lower indices
    same = torch.ones(6, C).to(cuda)
    assert ops.support_select(same, 2, 2, 3).cpu().tolist() == [[0, 1, 3], [0, 1, 2]]


# ---------------------------------------------------------------------------------------
# window bookkeeping kernels (csrc/window.cu), masked softmax, counted post-processing
# ---------------------------------------------------------------------------------------
def test_softmax_rows_masked(cuda):
    """hvr_softmax_rows_split_masked: blocks of `slot` keys of which only the first seg_counts[problem][block] take
    part == the softmax over the compacted key set (per row: same values at the live keys, exactly 0 elsewhere);
    with every count == slot the unmasked kernel's bits; a fully masked row gives zeros, not NaN."""
    from hvrnet_b200 import ops
    g = torch.Generator().manual_seed(3)
    V, nq, slot, segs = 3, 37, 300, 5
    cols = slot * segs
    S = (torch.randn(V * nq, cols + 44, generator=g) * 3).to(cuda)
    cnt = torch.tensor([[300, 17, 0, 299, 1], [300] * 5, [0, 0, 0, 0, 0]], dtype=torch.int32)
    Pm = ops.merge(ops.softmax_rows_split(S, cols, ld_p=cols + 44, seg_counts=cnt.to(cuda), slot=slot, rows_per_problem=nq)).cpu()
    Sc = S.cpu().double()
    for v in range(V):
        live = torch.cat([torch.arange(slot) < int(cnt[v, b]) for b in range(segs)])
        rows = slice(v * nq, (v + 1) * nq)
        assert not bool(Pm[rows, :cols][:, ~live].any()) and not bool(Pm[rows, cols:].any())
        if live.any():
            ref = torch.softmax(Sc[rows, :cols][:, live], 1)          # P is stored split-bf16: 2^-16 relative
            assert float(((Pm[rows, :cols][:, live].double() - ref).abs() / (ref + 1e-6)).max()) < 2e-5
        else:
            assert not bool(Pm[rows].any()) and bool(torch.isfinite(Pm[rows]).all())
    full = ops.softmax_rows_split(S, cols, ld_p=cols + 44)
    allc = torch.full((V, segs), slot, dtype=torch.int32, device=cuda)
    msk = ops.softmax_rows_split(S, cols, ld_p=cols + 44, seg_counts=allc, slot=slot, rows_per_problem=nq)
    assert torch.equal(full.hi.view(torch.int16), msk.hi.view(torch.int16))
    assert torch.equal(full.lo.view(torch.int16), msk.lo.view(torch.int16))
    # slot not a multiple of 4: the scalar path
    cnt2 = torch.tensor([[5, 0, 7]], dtype=torch.int32)
    S2 = torch.randn(4, 21, generator=g).to(cuda)
    P2 = ops.merge(ops.softmax_rows_split(S2, 21, ld_p=64, seg_counts=cnt2.to(cuda), slot=7, rows_per_problem=4)).cpu()
    live = torch.cat([torch.arange(7) < int(c) for c in cnt2[0]])
    ref2 = torch.softmax(S2.cpu().double()[:, live], 1)
    assert float(((P2[:, :21][:, live].double() - ref2).abs() / (ref2 + 1e-6)).max()) < 2e-5
    assert not bool(P2[:, :21][:, ~live].any())


def test_window_rois_and_gathers(cuda):
    """hvr_window_rois / hvr_gather_rows_split / hvr_support_index against their definitions (index arithmetic only)."""
    from hvrnet_b200 import ops
    g = torch.Generator().manual_seed(8)
    V, T, P, key = 3, 4, 8, 2
    F = V * T
    props = torch.rand(F, P, 5, generator=g) * 100
    counts = torch.randint(0, P + 1, (F,), generator=g, dtype=torch.int32)
    perm = torch.randperm(F, generator=g)
    rois, rk, seg, kc = ops.window_rois(props.to(cuda), counts.to(cuda), perm.to(cuda), V, T, key, n_segs=T + 2)
    Npad = ops.round_up(T * P, 64)
    rois, rk, seg, kc = rois.cpu().view(V, Npad, 5), rk.cpu().view(V, P, 5), seg.cpu(), kc.cpu()
    for v in range(V):
        for t in range(T):
            slot = int(perm[v * T + t])
            blk = rois[v, t * P:(t + 1) * P]
            assert torch.equal(blk[:, 1:], props[slot, :, :4]) and bool((blk[:, 0] == slot).all())
            assert int(seg[v, t]) == int(counts[slot])
        assert not bool(rois[v, T * P:].any()) and not bool(seg[v, T:].any())
        ks = int(perm[v * T + key])
        assert torch.equal(rk[v, :, 1:], props[ks, :, :4]) and not bool(rk[v, :, 0].any()) and int(kc[v]) == int(counts[ks])
    # identity perm
    r2 = ops.window_rois(props.to(cuda), counts.to(cuda), None, V, T, key)[0].cpu().view(V, Npad, 5)
    assert bool((r2[1, P:2 * P, 0] == T + 1).all())
    # gathers
    src = ops.split(torch.randn(40, 16, generator=g).to(cuda))
    dst = ops.Split.zeros((3 * 20, 16), cuda)
    ops.gather_rows(src, dst, 3, 5, src_rpp=10, src_row0=2, dst_rpp=20, dst_row0=7)
    for p in range(3):
        assert torch.equal(dst.hi[p * 20 + 7:p * 20 + 12], src.hi[p * 10 + 2:p * 10 + 7])
        assert torch.equal(dst.lo[p * 20 + 7:p * 20 + 12], src.lo[p * 10 + 2:p * 10 + 7])
    assert not bool(dst.hi[:7].any())
    idx = torch.tensor([3, -1, 39, 0, 0, 17], dtype=torch.int32, device=cuda)
    d2 = ops.Split.empty((6, 16), cuda)
    ops.gather_rows(src, d2, 1, 6, idx=idx)
    for j, i in enumerate(idx.tolist()):
        assert torch.equal(d2.hi[j], src.hi[i]) if i >= 0 else not bool(d2.hi[j].any())
    # support index: global key frame g of rank g // vpr
    sel = torch.tensor([[4, 1, -1], [0, 5, 2]], dtype=torch.int64, device=cuda)
    vpr, Pk, rank_stride, cstride = 2, 3, 50, 64
    pool_counts = torch.arange(3 * cstride, dtype=torch.int32, device=cuda)
    segc = torch.zeros((2, 7), dtype=torch.int32, device=cuda)
    ix = ops.support_index(sel, pool_counts, cstride, vpr, rank_stride, Pk, 4, segc).cpu()
    for v in range(2):
        for s_ in range(3):
            gk = int(sel[v, s_])
            want = [-1] * Pk if gk < 0 else [(gk // vpr) * rank_stride + (gk % vpr) * Pk + j for j in range(Pk)]
            assert ix[v, s_ * Pk:(s_ + 1) * Pk].tolist() == want
            assert int(segc[v, 4 + s_]) == (0 if gk < 0 else (gk // vpr) * cstride + gk % vpr)


@pytest.mark.parametrize('n_valid', [[300, 117, 1], [0, 300, 299]])
def test_det_postprocess_counted(cuda, n_valid):
    """hvr_det_postprocess_batched_ex: problem g with n_valid[g] proposals inside a block of n rows == the same call on
    exactly those n_valid[g] rows (labels, counts, boxes, scores identical); roi_idx = the row every detection came from
    (checked against the oracle's multiclass_nms rows)."""
    from hvrnet_b200 import ops
    from tests import parity_tools as PT
    g = torch.Generator().manual_seed(4)
    G, n = 3, 300
    x1, y1 = torch.rand(G * n, generator=g) * 700, torch.rand(G * n, generator=g) * 400
    rois = torch.stack([torch.zeros(G * n), x1, y1, x1 + 30 + torch.rand(G * n, generator=g) * 200,
                        y1 + 30 + torch.rand(G * n, generator=g) * 150], 1)
    cls = torch.randn(G * n, 31, generator=g) * 2
    reg = torch.randn(G * n, 4, generator=g) * 0.5
    nv = torch.tensor(n_valid, dtype=torch.int32)
    d, l, k, ridx = ops.det_postprocess_batched(rois.to(cuda), cls.to(cuda), reg.to(cuda), G, (600, 1000), 1.0, True,
                                                n_valid=nv.to(cuda), want_idx=True)
    for p in range(G):
        m = int(nv[p])
        kk = int(k[p])
        if m == 0:
            assert kk == 0
            continue
        sl = slice(p * n, p * n + m)
        t = PT.det_trace(rois[sl], cls[sl], reg[sl], (600, 1000), 1.0, True)
        assert kk == t['dets'].shape[0]
        assert l[p, :kk].cpu().tolist() == t['labels'].tolist() and ridx[p, :kk].cpu().tolist() == t['rows'].tolist()
        assert float((d[p, :kk, :4].cpu() - t['dets'][:, :4]).abs().max()) < 1e-3
        assert float((d[p, :kk, 4].cpu() - t['dets'][:, 4]).abs().max()) < 1e-6
        d1, l1, k1 = ops.det_postprocess(rois[sl].to(cuda), cls[sl].to(cuda), reg[sl].to(cuda), (600, 1000), 1.0, True)
        assert int(k1) == kk and torch.equal(d1[:kk], d[p, :kk]) and torch.equal(l1[:kk], l[p, :kk])


# ---------------------------------------------------------------------------------------
# C-ABI composites for non-Python hosts (csrc/relation.cu)
# ---------------------------------------------------------------------------------------
def test_c_host_relation_block(cuda, tmp_path):
    """tests/host/relation_host.c - a plain-C host: hvr_pack_linear + hvr_relation_fwd against its own double-precision
    restatement of forward_single_selsa (1e-3 relative), built with gcc and run here."""
    import subprocess
    from tests.test_host import _build_c_host
    exe = _build_c_host(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and 'OK' in r.stdout, (r.stdout, r.stderr)


def test_abi_packing_and_relation_equal_the_python_engine(cuda):
    """hvr_pack_conv_bn / hvr_pack_linear produce the bits engine.pack_* produce (fp64 BN fold, K-major re-layout,
    padding, split), and hvr_relation_fwd returns the bits engine.relation returns (the same six launches)."""
    import ctypes
    from hvrnet_b200 import _lib, engine, ops
    L = _lib.lib()
    g = torch.Generator().manual_seed(6)
    fp = lambda t: ctypes.c_void_p(t.data_ptr())
    # conv + BN (+ bias)
    for (co, ci, k, with_bn, with_bias) in ((96, 40, 3, True, False), (70, 64, 1, False, True), (64, 24, 3, True, True)):
        w = torch.randn(co, ci, k, k, generator=g)
        sd = {'c.weight': w}
        bn = [torch.rand(co, generator=g) + 0.5, torch.randn(co, generator=g), torch.randn(co, generator=g), torch.rand(co, generator=g) + 0.5]
        if with_bn:
            sd.update({'b.weight': bn[0], 'b.bias': bn[1], 'b.running_mean': bn[2], 'b.running_var': bn[3]})
        bias = torch.randn(co, generator=g) if with_bias else None
        if with_bias:
            sd['c.bias'] = bias
        ref = engine._pack_conv_bn(sd, 'c', 'b' if with_bn else None, cuda, bias_name='c.bias' if with_bias else None)
        R, C = L.hvr_packed_rows(co), L.hvr_packed_cols(k * k * ci)
        assert (R, C) == tuple(ref.w.shape)
        hi = torch.empty((R, C), dtype=torch.bfloat16, device=cuda)
        lo = torch.empty_like(hi)
        b = torch.empty(R, dtype=torch.float32, device=cuda)
        null = ctypes.c_void_p(0)
        rc = L.hvr_pack_conv_bn(fp(w), *( [fp(t) for t in bn] if with_bn else [null] * 4), 1e-5,
                                fp(bias) if with_bias else null, co, ci, k, k, fp(hi), fp(lo), fp(b), null)
        assert rc == 0
        assert torch.equal(hi.view(torch.int16), ref.w.hi.view(torch.int16)) and torch.equal(lo.view(torch.int16), ref.w.lo.view(torch.int16))
        assert torch.equal(b, ref.bias)
    # linear with a column permutation (fc_new_1's NHWC order)
    n, kk = 100, 4 * 3 * 3
    wl, bl = torch.randn(n, kk, generator=g), torch.randn(n, generator=g)
    perm = engine.nhwc_perm(4, 3)
    ref = engine.pack_linear(wl, bl, cuda, col_perm=perm)
    hi = torch.empty(tuple(ref.w.shape), dtype=torch.bfloat16, device=cuda)
    lo = torch.empty_like(hi)
    b = torch.empty(ref.w.shape[0], dtype=torch.float32, device=cuda)
    pi = perm.to(torch.int32).contiguous()
    assert L.hvr_pack_linear(fp(wl), fp(bl), n, kk, fp(pi), fp(hi), fp(lo), fp(b), ctypes.c_void_p(0)) == 0
    assert torch.equal(hi.view(torch.int16), ref.w.hi.view(torch.int16)) and torch.equal(lo.view(torch.int16), ref.w.lo.view(torch.int16))
    assert torch.equal(b, ref.bias)
    # one relation block: all-row queries and key-only queries
    D, N = 256, 450
    P = {}
    for name in ('q1', 'k1', 'o1'):
        P[name] = engine.pack_linear(torch.randn(D, D, generator=g) * 0.2, torch.randn(D, generator=g) * 0.1, cuda)
    X = ops.split(torch.randn(N, D, generator=g).to(cuda))
    XT = ops.transpose_split(X, D)
    W = _lib.HvrRelationWeights(D, *[x for nm in ('q1', 'k1', 'o1') for x in (P[nm].w.hi.data_ptr(), P[nm].w.lo.data_ptr(), P[nm].bias.data_ptr())])
    for (s0, nq) in ((0, N), (128, 100)):
        q_range = None if nq == N else (s0, nq)
        want = engine.relation(P, 1, X, XT, q_range=q_range, res=X[s0:s0 + nq])
        out = ops.Split.empty((nq, D), cuda)
        wsb = L.hvr_relation_workspace_bytes(nq, N, D)
        ws = torch.empty(wsb, dtype=torch.uint8, device=cuda)
        xq = X[s0:s0 + nq]
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = L.hvr_relation_fwd(ctypes.byref(W), fp(X.hi), fp(X.lo), D, N, fp(xq.hi) if q_range else None,
                                fp(xq.lo) if q_range else None, D, nq, fp(xq.hi), fp(xq.lo), D, 1, fp(out.hi), fp(out.lo), D,
                                fp(ws), wsb, st)
        assert rc == 0
        assert torch.equal(out.hi.view(torch.int16), want.hi.view(torch.int16))
        assert torch.equal(out.lo.view(torch.int16), want.lo.view(torch.int16))


def test_abi_layer_composites_equal_the_python_engine(cuda):
    """hvr_conv_fwd / hvr_linear_fwd (descriptor filling in C++) return the bits of engine.conv / engine.lin: dilated 3x3,
    strided 1x1 with residual, plain 1x1 with fp32 output, and a linear layer with residual + ReLU."""
    import ctypes
    from hvrnet_b200 import _lib, engine, ops
    L = _lib.lib()
    g = torch.Generator().manual_seed(12)
    fp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for (B, H, W, ci, co, k, dil, stride, use_res, f32) in ((2, 38, 63, 128, 96, 3, 2, 1, False, False),
                                                             (2, 38, 63, 256, 128, 1, 1, 2, True, False),
                                                             (1, 20, 24, 64, 60, 1, 1, 1, False, True)):
        x = ops.nchw_to_nhwc_split(torch.randn(B, ci, H, W, generator=g).to(cuda))
        w = torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5
        cp = engine.ConvP(engine.pack_conv(w, None, cuda), torch.randn(engine.round_up(co, 64), generator=g).to(cuda), co, k, ci, dil)
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        npad = cp.w.shape[0]
        res = ops.Split((torch.randn(B, Ho, Wo, npad, generator=g)).to(cuda).bfloat16(),
                        (torch.randn(B, Ho, Wo, npad, generator=g) * 0.01).to(cuda).bfloat16()) if use_res else None
        want, want_f = engine.conv(x, cp, stride=stride, relu=True, res=res, want_split=not f32, want_f32=f32)
        out = ops.Split.empty((B, Ho, Wo, npad), cuda) if not f32 else None
        of = torch.empty((B, Ho, Wo, engine.round_up(npad, 4)), dtype=torch.float32, device=cuda) if f32 else None
        rc = L.hvr_conv_fwd(fp(x.hi), fp(x.lo), B, H, W, ci, fp(cp.w.hi), fp(cp.w.lo), fp(cp.bias), co, k, dil, stride,
                            fp(res.hi) if res else None, fp(res.lo) if res else None, 1,
                            fp(out.hi) if out else None, fp(out.lo) if out else None, fp(of), st())
        assert rc == 0
        if f32:
            assert torch.equal(of, want_f)
        else:
            assert torch.equal(out.hi.view(torch.int16), want.hi.view(torch.int16))
            assert torch.equal(out.lo.view(torch.int16), want.lo.view(torch.int16))
    M, K, N = 700, 200, 130
    xs = ops.split(torch.randn(M, K, generator=g).to(cuda))
    lp = engine.pack_linear(torch.randn(N, K, generator=g) * 0.1, torch.randn(N, generator=g), cuda)
    npad = lp.w.shape[0]                                   # engine.lin computes the 64-padded rows (exact zeros)
    res = ops.split(torch.randn(M, npad, generator=g).to(cuda))
    want, want_f, _ = engine.lin(xs, lp, relu=True, res=res, want_f32=True)
    out = ops.Split.empty(tuple(want.shape), cuda)
    of = torch.empty(tuple(want_f.shape), dtype=torch.float32, device=cuda)
    rc = L.hvr_linear_fwd(fp(xs.hi), fp(xs.lo), M, K, xs.hi.stride(0), fp(lp.w.hi), fp(lp.w.lo), fp(lp.bias), npad,
                          fp(res.hi), fp(res.lo), res.hi.stride(0), 1, 1.0, fp(out.hi), fp(out.lo), out.hi.stride(0), fp(of),
                          of.stride(0), st())
    assert rc == 0
    assert torch.equal(out.hi.view(torch.int16), want.hi.view(torch.int16))
    assert torch.equal(out.lo.view(torch.int16), want.lo.view(torch.int16)) and torch.equal(of, want_f)
