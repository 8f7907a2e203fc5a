"""1x1 convolutions N=512 at growing batch (operand size vs the 126 MB L2) and K: isolates how the
pair kernel behaves when its A operand streams from HBM.  Usage: python scripts/k_sweep_bench.py [out.csv]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hvrnet_b200 import _lib, engine, ops  # noqa: E402

dev = torch.device('cuda:0')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(B, C, N, reps=6):
    x = ops.nchw_to_nhwc_split(torch.randn(B, C, 38, 63, device=dev))
    w = torch.randn(N, C, 1, 1) / math.sqrt(C)
    cp = engine.ConvP(engine.pack_conv(w, None, dev), torch.zeros(N, device=dev), N, 1, C, 1)
    ts = []
    for i in range(reps + 2):
        flush.zero_()                                  # cold L2 for every launch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.conv(x, cp, relu=True)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    return us, 2.0 * B * 38 * 63 * N * C / us / 1e6


out = open(sys.argv[1], 'w') if len(sys.argv) > 1 else sys.stdout
out.write('B,C,N,flags,us,algorithmic_TFLOPs\n')
for C in (512, 1024, 2048):
    for B in (15, 105):
        for flags in (0, 6 << 13, 12 << 13):
            _lib.lib().hvr_debug_force_bn(flags)
            us, tf = run(B, C, 512)
            _lib.lib().hvr_debug_force_bn(0)
            out.write('%d,%d,%d,%d,%.1f,%.1f\n' % (B, C, 512, flags, us, tf))
            out.flush()
