set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gpu_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2a_gpu_tests.log
timeout 900 python tests/parity_report.py > gpurun_out/r2a_parity_report.txt 2> gpurun_out/r2a_parity_report.err; echo "parity rc=$?"
tail -30 gpurun_out/r2a_parity_report.txt
timeout 300 python scripts/roi_bench.py > gpurun_out/r2a_roi_bench.txt 2>&1; echo "roi rc=$?"
cat gpurun_out/r2a_roi_bench.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
cat gpurun_out/r2a_bench.json
