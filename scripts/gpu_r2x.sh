#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/layer3_bench.py > gpurun_out/r2x_layer3_bench.txt 2>&1
cat gpurun_out/r2x_layer3_bench.txt
