"""Trunk layer3 at the bench's batch (7 key frames, 38 x 63 x 1024): the three convolutions of one bottleneck block, isolated
launches (CUDA events, 20 back-to-back) and the whole 23-block layer as the pipeline runs it (engine.res_layer, graph replay),
for the tile variants hvr_debug_force_bn selects.      python scripts/layer3_bench.py"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hvrnet_b200 import _lib, engine, ops  # noqa: E402

dev = torch.device('cuda:0')
B, H, W = 7, 38, 63
L = _lib.lib()


def time_fn(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def mk(C, N, k, res=False):
    x = ops.nchw_to_nhwc_split(torch.randn(B, C, H, W, device=dev))
    w = torch.randn(N, C, k, k) / math.sqrt(C * k * k)
    cp = engine.ConvP(engine.pack_conv(w, None, dev), torch.zeros(max(N, 64), device=dev), N, k, C, 1)
    r = ops.nchw_to_nhwc_split(torch.randn(B, N, H, W, device=dev)) if res else None
    return (lambda: engine.conv(x, cp, relu=True, res=r)), 2.0 * B * H * W * N * C * k * k


cases = [('conv1 1x1 1024->256', mk(1024, 256, 1)), ('conv2 3x3 256->256', mk(256, 256, 3)),
         ('conv3 1x1 256->1024 + res', mk(256, 1024, 1, True))]
print('case, variant, us, algorithmic TFLOP/s   (M = %d rows)' % (B * H * W))
for name, (fn, fl) in cases:
    for label, flag in (('auto', 0), ('single-CTA BN=128', 128), ('single-CTA BN=256', 256), ('pair 256x256', 512), ('pair 256x128', 640)):
        L.hvr_debug_force_bn(flag)
        try:
            us = time_fn(fn)
        except Exception as e:                                # noqa: BLE001
            print('%s, %s, failed: %s' % (name, label, e))
            continue
        finally:
            L.hvr_debug_force_bn(0)
        print('%s, %s, %.1f, %.1f' % (name, label, us, fl / us / 1e6))
