"""Trunk layer3 (23 bottleneck blocks, 69 GEMM launches, 7 key frames: M = 16 758 rows after the first block) as the pipeline
runs it: engine.res_layer captured in a CUDA graph, replayed; per-launch average for the tile / epilogue variants.
    python scripts/layer3_chain_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hvrnet_b200 import _lib, configs, engine, ops  # noqa: E402

dev = torch.device('cuda:0')
L = _lib.lib()
m, sd, w = configs.build_workload('hrnmp', dev)
P = engine.pack_trunk(sd, dev)
blocks = P['layers'][2]
x = ops.nchw_to_nhwc_split(torch.randn(7, 512, 76, 126, device=dev) * 0.5)
fl = 0.0


def run():
    return engine.res_layer(x, blocks, 2)


for label, flag in (('heuristic (shipped)', 0), ('never deep epilogue', 4096), ('always deep epilogue', 2048),
                    ('no lean variant', 1 << 17), ('CLC tile hand-out', 1 << 18), ('per-row epilogue', 1024)):
    L.hvr_debug_force_bn(flag)
    try:
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                out = run()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / 10 * 1e3
        print('%-24s layer3 = %.1f us per replay, %.1f us per launch (69 launches)' % (label, us, us / 69))
    finally:
        L.hvr_debug_force_bn(0)
