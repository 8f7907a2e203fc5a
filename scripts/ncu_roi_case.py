"""One batched RoIAlign launch sequence (bench.py's shape: 105 frames x 300 RoIs, NHWC in -> split rows out) for
an ncu capture:
  ncu --set full --clock-control none --import-source on -k regex:roi_align -s 3 -c 1 -o out \\
      python scripts/ncu_roi_case.py [variant]        # hvr_debug_roi_variant: 0 = strict heuristic (sn2), 1 = per-bin kernel,
                                                      # 5 = fast slab kernel, 6 = fast RoI-per-CTA kernel
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hvrnet_b200 import _lib, ops  # noqa: E402

dev = torch.device('cuda:0')
VARIANT = int(sys.argv[1]) if len(sys.argv) > 1 else 0
_lib.lib().hvr_debug_roi_variant(VARIANT)
ARITH = 'fast' if VARIANT >= 4 else 'strict'
g = torch.Generator().manual_seed(5)
Tn = 105
feat = torch.randn(Tn, 38, 63, 256, generator=g).to(dev)
x1 = torch.rand(Tn * 300, generator=g) * 800
y1 = torch.rand(Tn * 300, generator=g) * 450
wh = torch.rand(Tn * 300, 2, generator=g) * 380 + 16
rois = torch.stack([(torch.arange(Tn * 300) // 300).float(), x1, y1, (x1 + wh[:, 0]).clamp(max=999),
                    (y1 + wh[:, 1]).clamp(max=599)], 1).to(dev)
for _ in range(5):
    ops.roi_align(feat, rois, feat_nhwc=True, out_nhwc=True, want_split=True, want_f32=False, arithmetic=ARITH)
torch.cuda.synchronize()
print('done')
