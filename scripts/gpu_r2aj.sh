#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "roi_align" > gpurun_out/r2aj_tests.txt 2>&1; tail -3 gpurun_out/r2aj_tests.txt
timeout 300 python scripts/roi_bench.py 2>&1 | grep "T=105 side 16-396" | grep "shipped"
python __graft_entry__.py smoke 2>&1 | tail -1
