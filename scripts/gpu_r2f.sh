set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "roi_align or softmax or window_rois" > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2f_tests.log
timeout 300 python scripts/roi_bench.py > gpurun_out/r2f_roi_bench.txt 2>&1; echo "roi rc=$?"
grep -v reference_kernel gpurun_out/r2f_roi_bench.txt
timeout 300 python scripts/relation_stage_bench.py > gpurun_out/r2f_relation_stage_bench.txt 2>&1; grep -i "softmax" gpurun_out/r2f_relation_stage_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:roi_align_sep -s 2 -c 1 -o gpurun_out/r2f_roi_sep python scripts/ncu_roi_case.py 6 > gpurun_out/r2f_ncu_roi.log 2>&1; echo "ncu roi rc=$?"
