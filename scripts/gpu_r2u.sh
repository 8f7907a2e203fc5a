#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the fast RoIAlign variants and the new layer composites
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "roi_align_fast_variant_vs_strict or abi_layer" > gpurun_out/r2u_memcheck.txt 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2u_memcheck.txt | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "roi_align_fast_variant_vs_strict" > gpurun_out/r2u_racecheck.txt 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2u_racecheck.txt | tail -3
