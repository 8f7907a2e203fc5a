set -x
mkdir -p gpurun_out
timeout 900 python tests/parity_report.py > gpurun_out/r2o_parity_report.txt 2> gpurun_out/r2o_parity_report.err; echo "parity rc=$?"
grep -c "near-tie" gpurun_out/r2o_parity_report.txt; grep -c "NOT EXPLAINED" gpurun_out/r2o_parity_report.txt; grep "replay" gpurun_out/r2o_parity_report.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-streaming --no-other-workloads --gemm-report gpurun_out/r2o_gemm_shapes.csv > gpurun_out/r2o_bench.json 2>/dev/null; echo "bench rc=$?"
cat gpurun_out/r2o_gemm_shapes.csv
