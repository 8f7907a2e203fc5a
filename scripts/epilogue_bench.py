"""Micro-benchmark of the pair kernel's two epilogues on the residual / short-K layers of the hot path
(isolated launches, CUDA events): 3-stage ring + 32-column staging (flag 4096) against the deep
epilogue (2-stage ring, 3 in-place 64-column buffers per warp, flag 2048).
Usage: python scripts/epilogue_bench.py [out.csv]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hvrnet_b200 import _lib, engine, ops  # noqa: E402

dev = torch.device('cuda:0')
REPS = 20


def time_fn(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REPS * 1e3     # us


def conv_case(B, H, W, C, N, k, dil, res):
    x = ops.nchw_to_nhwc_split(torch.randn(B, C, H, W, device=dev))
    w = torch.randn(N, C, k, k) / math.sqrt(C * k * k)
    cp = engine.ConvP(engine.pack_conv(w, None, dev), torch.zeros(max(N, 64), device=dev), N, k, C, dil)
    r = ops.nchw_to_nhwc_split(torch.randn(B, N, H, W, device=dev)) if res else None
    flops = 2.0 * B * H * W * N * C * k * k
    mbytes = 4.0 * B * H * W * (C + N * (2 if res else 1)) / 1e6
    return (lambda: engine.conv(x, cp, relu=True, res=r)), flops, mbytes


def lin_case(M, K, N, res):
    a = ops.split(torch.randn(M, K, device=dev))
    w = ops.split(torch.randn(max(N, 64), K, device=dev) / math.sqrt(K))
    r = ops.split(torch.randn(M, N, device=dev)) if res else None
    return (lambda: ops.linear(a, w, N, relu=True, res=r)), 2.0 * M * K * N, 4.0 * M * (K + N * (2 if res else 1)) / 1e6


CASES = [
    ('trunk l1 conv3 1x1 64->256 +res B7', lambda: conv_case(7, 152, 252, 64, 256, 1, 1, True)),
    ('trunk l2 conv3 1x1 128->512 +res B7', lambda: conv_case(7, 76, 126, 128, 512, 1, 1, True)),
    ('trunk l3 conv3 1x1 256->1024 +res B7', lambda: conv_case(7, 38, 63, 256, 1024, 1, 1, True)),
    ('trunk l3 conv1 1x1 1024->256 B7', lambda: conv_case(7, 38, 63, 1024, 256, 1, 1, False)),
    ('trunk l3 conv2 3x3 256->256 B7', lambda: conv_case(7, 38, 63, 256, 256, 3, 1, False)),
    ('trunk l2 conv2 3x3 128->128 B7', lambda: conv_case(7, 76, 126, 128, 128, 3, 1, False)),
    ('trunk l2 conv1 1x1 512->128 B7', lambda: conv_case(7, 76, 126, 512, 128, 1, 1, False)),
    ('c5 conv3 1x1 512->2048 +res B15', lambda: conv_case(15, 38, 63, 512, 2048, 1, 1, True)),
    ('c5 conv1 1x1 2048->512 B15', lambda: conv_case(15, 38, 63, 2048, 512, 1, 1, False)),
    ('c5 conv2 3x3d2 512->512 B15', lambda: conv_case(15, 38, 63, 512, 512, 3, 2, False)),
    ('c5 down 1x1 1024->2048 B15', lambda: conv_case(15, 38, 63, 1024, 2048, 1, 1, False)),
    ('head out-proj 31808x1024x1024 +res', lambda: lin_case(31808, 1024, 1024, True)),
    ('head q/k 31808x1024x1024', lambda: lin_case(31808, 1024, 1024, False)),
    ('head out-proj 4500x1024x1024 +res', lambda: lin_case(4500, 1024, 1024, True)),
]

out = open(sys.argv[1], 'w') if len(sys.argv) > 1 else sys.stdout
out.write('case,epilogue,us,algorithmic_TFLOPs,min_traffic_GBps\n')
for name, mk in CASES:
    fn, flops, mbytes = mk()
    for label, flag in (('3stage+32col', 512 | 4096), ('deep', 512 | 2048), ('pair128', 640), ('auto', 0)):
        _lib.lib().hvr_debug_force_bn(flag)
        try:
            us = time_fn(fn)
        finally:
            _lib.lib().hvr_debug_force_bn(0)
        out.write('%s,%s,%.1f,%.1f,%.0f\n' % (name, label, us, flops / us / 1e6, mbytes / us * 1e3))
        out.flush()
