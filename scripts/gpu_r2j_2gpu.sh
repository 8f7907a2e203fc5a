set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err; echo "bench2 rc=$?"
tail -5 gpurun_out/r2j_bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench_2gpu.json'))
print('value', d['value'], 'e2e', d['e2e']['value'])
print(json.dumps(d.get('inter_video'), indent=1))
print('streaming', d['streaming']['value'], d['streaming']['e2e']['value'])
PY
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "faster_rcnn or two_ranks" > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2j_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-streaming > gpurun_out/r2j_bench_1gpu.json 2> gpurun_out/r2j_bench_1gpu.err; echo "bench1 rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2j_bench_1gpu.json'))
print(d['value']); print(json.dumps(d['other_workloads'], indent=1))"
