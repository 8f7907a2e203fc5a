#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "roi_align" > gpurun_out/r2r_tests.txt 2>&1
tail -3 gpurun_out/r2r_tests.txt
timeout 600 python scripts/roi_bench.py > gpurun_out/r2r_roi_bench.txt 2>&1
grep "T=105" gpurun_out/r2r_roi_bench.txt | grep "row_program"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_align -s 3 -c 1 -f -o gpurun_out/r2r_roi_sepp4 \
    python scripts/ncu_roi_case.py 6 > gpurun_out/r2r_ncu.log 2>&1
tail -2 gpurun_out/r2r_ncu.log
