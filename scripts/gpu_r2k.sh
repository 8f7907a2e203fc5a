set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "faster_rcnn or graphs_follow" > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2k_tests.log
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  HVR_FORK_POST=$1 HVR_FORK_PROPOSALS=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-streaming --no-other-workloads > gpurun_out/r2k_bench_$1$2.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2k_bench_$1$2.json'))
print('fork_post=$1 fork_proposals=$2 V=7: value %.2f ms/step %.3f e2e %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
for cfg in "1 1" "0 1" "0 0"; do
  set -- $cfg
  HVR_FORK_POST=$1 HVR_FORK_PROPOSALS=$2 timeout 600 python bench.py --steps 5 --warmup 3 --videos-per-gpu 32 --no-cpu-baseline --no-streaming --no-other-workloads > gpurun_out/r2k_bench_v32_$1$2.json 2>/dev/null
  python -c "
import json
d=json.load(open('gpurun_out/r2k_bench_v32_$1$2.json'))
print('fork_post=$1 fork_proposals=$2 V=32: value %.2f ms/step %.3f' % (d['value'], d['ms_per_step']))"
done
