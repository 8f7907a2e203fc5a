set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2aa_bench_2gpu.json 2> gpurun_out/r2aa_bench_2gpu.err; echo "bench2 rc=$?"
tail -3 gpurun_out/r2aa_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2aa_bench_2gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
print(json.dumps(d.get('inter_video'), indent=1)[:1500])
print('streaming', d['streaming']['value'], d['streaming']['e2e']['value'])
PY
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "two_ranks" > gpurun_out/r2aa_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2aa_tests.log
