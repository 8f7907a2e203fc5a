#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/layer3_chain_bench.py > gpurun_out/r2x_layer3_chain.txt 2>&1
cat gpurun_out/r2x_layer3_chain.txt | tail -12
