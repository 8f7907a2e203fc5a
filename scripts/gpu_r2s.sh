set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2s_smoke.log
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2s_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r2s_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"
cat gpurun_out/r2s_bench.json | cut -c1-1800; tail -3 gpurun_out/r2s_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2s_bench_reference.json 2> gpurun_out/r2s_bench_reference.err; echo "ref rc=$?"; cat gpurun_out/r2s_bench_reference.json | cut -c1-600
HVR_NO_PREFETCH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 1500 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-streaming --no-other-workloads > gpurun_out/r2s_ncu_bench.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_launches.py gpurun_out/r2s_launches.csv > gpurun_out/r2s_launches.txt 2>&1; grep -c "at::" gpurun_out/r2s_launches.txt; head -3 gpurun_out/r2s_launches.txt
