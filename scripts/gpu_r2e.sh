set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2e_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/r2e_gpu_tests.log
timeout 300 python scripts/relation_stage_bench.py > gpurun_out/r2e_relation_stage_bench.txt 2>&1; echo "rel rc=$?"
cat gpurun_out/r2e_relation_stage_bench.txt
HVR_NO_PREFETCH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 1500 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-streaming --no-other-workloads > gpurun_out/r2e_ncu_bench.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_launches.py gpurun_out/r2e_launches.csv > gpurun_out/r2e_launches.txt 2>&1; head -60 gpurun_out/r2e_launches.txt
HVR_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:igemm --csv --log-file gpurun_out/r2e_igemm_metrics.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-streaming --no-other-workloads > gpurun_out/r2e_ncu_range.log 2>&1; echo "ncu range rc=$?"
python scripts/ncu_metrics_summary.py gpurun_out/r2e_igemm_metrics.csv 2 gpurun_out/r2e_igemm_ncu_step.json > gpurun_out/r2e_igemm_metrics.txt 2>&1; cat gpurun_out/r2e_igemm_metrics.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_align_sep -s 2 -c 1 -o gpurun_out/r2e_roi_sep python scripts/ncu_roi_case.py 6 > gpurun_out/r2e_ncu_roi.log 2>&1; echo "ncu roi rc=$?"
