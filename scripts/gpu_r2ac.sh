#!/bin/bash
mkdir -p gpurun_out
timeout 800 python scripts/inter_step_times.py 7 30 > gpurun_out/r2ad_step_times_v7_pool.txt 2>&1
grep "prefetch=\|Error\|error" gpurun_out/r2ad_step_times_v7_pool.txt | cut -c1-250
grep "host phases" gpurun_out/r2ad_step_times_v7_pool.txt | grep -v "step 29" | cut -c60-330
timeout 900 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/r2ad_pipeline_tests.txt 2>&1; tail -3 gpurun_out/r2ad_pipeline_tests.txt
