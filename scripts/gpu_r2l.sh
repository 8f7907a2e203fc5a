set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "roi_align" > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2l_tests.log
timeout 300 python scripts/roi_bench.py > gpurun_out/r2l_roi_bench.txt 2>&1; echo "roi rc=$?"
grep -v "reference_kernel\|generic\|slab" gpurun_out/r2l_roi_bench.txt
