#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-streaming --no-other-workloads --gemm-report gpurun_out/r2y_gemm_shapes.csv > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
tail -2 gpurun_out/r2y_bench.err
FORK=0; HVR_FORK_PROPOSALS=0 HVR_FORK_POST=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-streaming --no-other-workloads --gemm-report gpurun_out/r2y_gemm_shapes_nofork.csv > gpurun_out/r2y_bench_nofork.json 2> gpurun_out/r2y_bench_nofork.err
awk -F, '$6>400' gpurun_out/r2y_gemm_shapes_timeline.csv
echo ---- ; awk -F, '$6>400' gpurun_out/r2y_gemm_shapes_nofork_timeline.csv
