"""Runs one residual 1x1 convolution of the hot path in isolation for an ncu capture:
  ncu --set full --clock-control none --import-source on -k regex:igemm_tc2 -s 3 -c 1 -o out \
      python scripts/ncu_case.py c5conv3|l3conv3 deep|3stage
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hvrnet_b200 import _lib, engine, ops  # noqa: E402

dev = torch.device('cuda:0')
case = sys.argv[1] if len(sys.argv) > 1 else 'c5conv3'
mode = sys.argv[2] if len(sys.argv) > 2 else 'deep'
B, H, W, C, N = {'c5conv3': (15, 38, 63, 512, 2048), 'l3conv3': (7, 38, 63, 256, 1024),
                 'c5conv1': (105, 38, 63, 2048, 512), 'c5conv1k1024': (105, 38, 63, 1024, 512)}[case]
x = ops.nchw_to_nhwc_split(torch.randn(B, C, H, W, device=dev))
r = ops.nchw_to_nhwc_split(torch.randn(B, N, H, W, device=dev)) if 'conv3' in case else None if 'conv3' in case else None
w = torch.randn(N, C, 1, 1) / math.sqrt(C)
cp = engine.ConvP(engine.pack_conv(w, None, dev), torch.zeros(N, device=dev), N, 1, C, 1)
_lib.lib().hvr_debug_force_bn(512 | (2048 if mode == 'deep' else 4096))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(6):
    flush.zero_()
    engine.conv(x, cp, relu=True, res=r)
torch.cuda.synchronize()
print('done: %s %s %.1f GFLOP per launch' % (case, mode, 2.0 * B * H * W * N * C / 1e9))
