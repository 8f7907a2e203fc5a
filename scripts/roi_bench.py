"""RoIAlign micro-benchmark (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hvrnet_b200 import _lib, ops  # noqa: E402

dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(5)
# (frames, RoI side range in pixels): bench.py's distribution at three batch sizes, then small / large RoIs only
for T, lo, hi in ((1, 16, 396), (15, 16, 396), (105, 16, 396), (105, 16, 112), (105, 224, 396)):
    feat = torch.randn(T, 38, 63, 256, generator=g).to(dev)
    n = T * 300
    x1 = torch.rand(n, generator=g) * 800
    y1 = torch.rand(n, generator=g) * 450
    wh = torch.rand(n, 2, generator=g) * (hi - lo) + lo
    rois = torch.stack([(torch.arange(n) // 300).float(), x1, y1, (x1 + wh[:, 0]).clamp(max=999),
                        (y1 + wh[:, 1]).clamp(max=599)], 1).to(dev)
    # generic = per-bin kernel (16 loads per output vector), sn2 = strict tap-reuse kernel, fast_roi_cta / fast_slab = the
    # separable FMA evaluation with one CTA per RoI / with CTA = (frame, 16-channel slab of the map in shared memory)
    for name, variant, arith in (("generic", 1, "strict"), ("sn2", 0, "strict"), ("fast_roi_cta_4ch", 6, "fast"),
                                 ("fast_roi_cta_8ch_adjacent", 107, "fast"), ("fast_roi_cta_8ch_interleaved", 117, "fast"),
                                 ("fast_roi_cta_16ch_interleaved", 127, "fast"), ("fast_roi_cta_8ch_interleaved_row_program [shipped]", 137, "fast"),
                                 ("fast_roi_cta_8ch_interleaved_row_program_prefetch", 147, "fast"),
                                 ("fast_roi_cta_8ch_interleaved_row_program_3ctas_80regs", 157, "fast"), ("fast_roi_cta_3ctas", 3, "fast"), ("fast_slab", 5, "fast")):
        _lib.lib().hvr_debug_roi_variant(0)
        _lib.lib().hvr_debug_roi_variant(2)
        _lib.lib().hvr_debug_roi_variant(8)
        _lib.lib().hvr_debug_roi_variant(13)
        if variant > 100:                                     # 1xy: channel layout 1x of the 8-channel kernel
            _lib.lib().hvr_debug_roi_variant(variant // 10)
            variant = 7
        _lib.lib().hvr_debug_roi_variant(6 if variant in (3, 7) else 4)
        _lib.lib().hvr_debug_roi_variant(variant)
        fn = lambda: ops.roi_align(feat, rois, feat_nhwc=True, out_nhwc=True, want_split=True, want_f32=False,
                                   arithmetic=arith)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / 10 * 1e3
        print('T=%d side %d-%d px variant=%s %.1f us  %.0f GB/s (algorithmic 17.51 MB/frame)' % (T, lo, hi, name, us, 17510256.0 * T / us / 1e3))
    for v_ in (0, 2, 7, 4, 13):
        _lib.lib().hvr_debug_roi_variant(v_)
    # the reference's own CUDA op (oracle/_ref, compiled unmodified for sm_100a): NCHW map in, NCHW fp32 out
    try:
        from oracle import build as obuild
        ref = obuild.load_ref_roi_align(True)
    except Exception:                                         # noqa: BLE001
        ref = None
    if ref is not None:
        fc = feat.permute(0, 3, 1, 2).contiguous()
        out = fc.new_zeros(n, 256, 7, 7)
        for _ in range(2):
            ref.forward(fc, rois, 7, 7, 1 / 16., 2, out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            ref.forward(fc, rois, 7, 7, 1 / 16., 2, out)
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / 5 * 1e3
        print('T=%d side %d-%d px variant=reference_kernel %.1f us  %.0f GB/s' % (T, lo, hi, us, 17510256.0 * T / us / 1e3))
