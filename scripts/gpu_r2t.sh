#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -x -q -m gpu -k "im2col or stem or trunk" > gpurun_out/r2t_tests.txt 2>&1
tail -3 gpurun_out/r2t_tests.txt
python - <<'PY'
import torch
from hvrnet_b200 import ops
x = torch.randn(7, 3, 608, 1008, device='cuda')
for _ in range(3): ops.im2col_stem(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): ops.im2col_stem(x)
b.record(); torch.cuda.synchronize()
print('im2col_stem 7 frames: %.1f us' % (a.elapsed_time(b) / 20 * 1e3))
PY
