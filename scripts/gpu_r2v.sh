#!/bin/bash
# compute-sanitizer memcheck over the whole kernel test file
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q -m gpu > gpurun_out/r2v_memcheck_kernels.txt 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|at 0x" gpurun_out/r2v_memcheck_kernels.txt | head -20
