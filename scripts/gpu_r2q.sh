#!/bin/bash
# ncu --set full of the shipped (lane-interleaved 8-channel) fast RoIAlign kernel on bench.py's launch
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_align -s 3 -c 1 -f -o gpurun_out/r2q_roi_sepp \
    python scripts/ncu_roi_case.py 6 > gpurun_out/r2q_ncu.log 2>&1
tail -3 gpurun_out/r2q_ncu.log
ls -la gpurun_out/r2q_roi_sepp.ncu-rep
