set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -30 gpurun_out/r2c_gpu_tests.log
timeout 300 python scripts/roi_bench.py > gpurun_out/r2c_roi_bench.txt 2>&1; echo "roi rc=$?"
cat gpurun_out/r2c_roi_bench.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-workloads > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
cat gpurun_out/r2c_bench.json; tail -5 gpurun_out/r2c_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 2500 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-streaming --no-other-workloads > gpurun_out/r2c_ncu_bench.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_launches.py gpurun_out/r2c_launches.csv > gpurun_out/r2c_launches.txt 2>&1; head -70 gpurun_out/r2c_launches.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_align_slab -s 2 -c 1 -o gpurun_out/r2c_roi_slab python scripts/ncu_roi_case.py 5 > gpurun_out/r2c_ncu_roi.log 2>&1; echo "ncu roi rc=$?"
