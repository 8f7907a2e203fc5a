#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k "eager_reissue or prefetch or graph_runner_matches" > gpurun_out/r2ae_tests.txt 2>&1; tail -5 gpurun_out/r2ae_tests.txt
