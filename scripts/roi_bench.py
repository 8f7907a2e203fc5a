"""RoIAlign micro-benchmark (CUDA events)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hvrnet_b200 import ops  # noqa: E402

dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(5)
for T in (1, 15, 64):
    feat = torch.randn(T, 38, 63, 256, generator=g).to(dev)
    n = T * 300
    x1 = torch.rand(n, generator=g) * 800
    y1 = torch.rand(n, generator=g) * 450
    wh = torch.rand(n, 2, generator=g) * 380 + 16
    rois = torch.stack([(torch.arange(n) // 300).float(), x1, y1, (x1 + wh[:, 0]).clamp(max=999),
                        (y1 + wh[:, 1]).clamp(max=599)], 1).to(dev)
    for resident in (False,):
        fn = lambda: ops.roi_align(feat, rois, feat_nhwc=True, out_nhwc=True, want_split=True, want_f32=False)
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
        print('T=%d resident=%s %.1f us  %.0f GB/s (algorithmic 17.51 MB/frame)' % (T, resident, us, 17510256.0 * T / us / 1e3))
