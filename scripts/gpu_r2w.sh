#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "window or softmax or nms or det or proposal or gather or support or transpose or im2col or maxpool" > gpurun_out/r2w_racecheck.txt 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2w_racecheck.txt | tail -3; grep -c "hazard" gpurun_out/r2w_racecheck.txt
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -x -k "ragged or faster_rcnn or stream or inter" > gpurun_out/r2w_memcheck_pipeline.txt 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2w_memcheck_pipeline.txt | tail -3
