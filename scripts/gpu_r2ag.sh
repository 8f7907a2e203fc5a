#!/bin/bash
# ncu range capture of the igemm launches of the timed region (final tree) -> profiles/igemm_ncu_step.json, and the per-shape table
mkdir -p gpurun_out
HVR_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:igemm --csv --log-file gpurun_out/r2ag_igemm_metrics.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-streaming --no-other-workloads > gpurun_out/r2ag_ncu_range.log 2>&1; echo "ncu range rc=$?"
python scripts/ncu_metrics_summary.py gpurun_out/r2ag_igemm_metrics.csv 2 gpurun_out/r2ag_igemm_ncu_step.json > gpurun_out/r2ag_igemm_metrics.txt 2>&1; cat gpurun_out/r2ag_igemm_metrics.txt
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-streaming --no-other-workloads --gemm-report gpurun_out/r2ag_gemm_shapes.csv > gpurun_out/r2ag_bench.json 2> gpurun_out/r2ag_bench.err; echo "bench rc=$?"
