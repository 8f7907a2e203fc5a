#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r2ak_gpu_tests.txt 2>&1; tail -3 gpurun_out/r2ak_gpu_tests.txt
python __graft_entry__.py smoke 2>&1 | tail -1
