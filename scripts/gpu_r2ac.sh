#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 600 python scripts/inter_loss_1gpu.py 32 8 2>&1 | tail -1 | cut -c1-420; done
