"""Runs the top kernel of the bench in isolation (the C5 3x3 dilated conv, 15 frames:
M = 35910, N = 512, K = 4608, 3-product split-bf16, CTA-pair tcgen05 kernel) for an ncu capture:
  ncu --set full --clock-control none --import-source on -k regex:igemm_tc2 -s 3 -c 2 -o out python scripts/ncu_top_kernel.py
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hvrnet_b200 import engine, ops  # noqa: E402

dev = torch.device('cuda:0')
B, H, W, C, N, k, dil = 15, 38, 63, 512, 512, 3, 2
x = ops.nchw_to_nhwc_split(torch.randn(B, C, H, W, device=dev))
w = torch.randn(N, C, k, k) / math.sqrt(C * k * k)
cp = engine.ConvP(engine.pack_conv(w, None, dev), torch.zeros(N, device=dev), N, k, C, dil)
for _ in range(6):
    engine.conv(x, cp, relu=True)
torch.cuda.synchronize()
print('done: %.1f GFLOP per launch' % (2.0 * B * H * W * N * C * k * k / 1e9))
