"""Micro-benchmark of hvr_igemm on the hot-path shapes (isolated launches, CUDA events).
Usage: python scripts/gemm_bench.py [out.csv]"""
import math
import sys
import os

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


def conv_case(B, H, W, C, N, k, dil, passes=3):
    x = ops.nchw_to_nhwc_split(torch.randn(B, C, H, W, device=dev))
    w = torch.randn(N, C, k, k) / math.sqrt(C * k * k)
    cp = engine.ConvP(engine.pack_conv(w, None, dev), torch.zeros(max(N, 64), device=dev), N, k, C, dil)
    flops = 2.0 * B * H * W * N * C * k * k
    return (lambda: engine.conv(x, cp, relu=True, passes=passes)), flops


def lin_case(M, K, N, passes=3):
    a = ops.split(torch.randn(M, K, device=dev))
    w = ops.split(torch.randn(max(N, 64), K, device=dev) / math.sqrt(K))
    return (lambda: ops.linear(a, w, N, passes=passes)), 2.0 * M * K * N


CASES = [
    ('trunk l3 3x3 256->256 B1', lambda p: conv_case(1, 38, 63, 256, 256, 3, 1, p)),
    ('trunk l3 1x1 256->1024 B1', lambda p: conv_case(1, 38, 63, 256, 1024, 1, 1, p)),
    ('trunk l3 1x1 1024->256 B1', lambda p: conv_case(1, 38, 63, 1024, 256, 1, 1, p)),
    ('trunk l1 1x1 64->256 B1', lambda p: conv_case(1, 152, 252, 64, 256, 1, 1, p)),
    ('c5 3x3d2 512->512 B15', lambda p: conv_case(15, 38, 63, 512, 512, 3, 2, p)),
    ('c5 1x1 512->2048 B15', lambda p: conv_case(15, 38, 63, 512, 2048, 1, 1, p)),
    ('c5 1x1 2048->512 B15', lambda p: conv_case(15, 38, 63, 2048, 512, 1, 1, p)),
    ('head 4500x1024x1024', lambda p: lin_case(4500, 1024, 1024, p)),
    ('head fc1 4500x12544x1024', lambda p: lin_case(4500, 12544, 1024, p)),
    ('lin 2394x2304x256', lambda p: lin_case(2394, 2304, 256, p)),
]

out = open(sys.argv[1], 'w') if len(sys.argv) > 1 else sys.stdout
out.write('case,bn,passes,us,algorithmic_TFLOPs\n')
for name, mk in CASES:
    for passes in (3, 1):
        for bn in (64, 128, 256, 512):
            if bn == 512 and passes != 3:
                continue
            fn, flops = mk(passes)
            _lib.lib().hvr_debug_force_bn(bn)
            try:
                us = time_fn(fn)
            finally:
                _lib.lib().hvr_debug_force_bn(0)
            out.write('%s,%d,%d,%.1f,%.1f\n' % (name, bn, passes, us, flops / us / 1e6))
            out.flush()
