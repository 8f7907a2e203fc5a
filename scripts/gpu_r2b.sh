set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2b_smoke.log
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2b_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -40 gpurun_out/r2b_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
cat gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-streaming --no-other-workloads > gpurun_out/r2b_ncu_bench.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_launches.py gpurun_out/r2b_launches.csv > gpurun_out/r2b_launches.txt 2>&1; head -70 gpurun_out/r2b_launches.txt
