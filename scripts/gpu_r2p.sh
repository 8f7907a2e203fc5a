#!/bin/bash
# interleaved-channel RoIAlign + new ABI composites
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "roi_align or abi" > gpurun_out/r2p_tests.txt 2>&1
tail -5 gpurun_out/r2p_tests.txt
timeout 600 python scripts/roi_bench.py > gpurun_out/r2p_roi_bench.txt 2>&1
grep "T=105" gpurun_out/r2p_roi_bench.txt
