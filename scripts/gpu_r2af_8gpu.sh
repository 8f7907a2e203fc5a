set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2af_bench_8gpu.json 2> gpurun_out/r2af_bench_8gpu.err; echo "bench8 rc=$?"
cut -c1-700 gpurun_out/r2af_bench_8gpu.json; tail -5 gpurun_out/r2af_bench_8gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2af_bench_8gpu.json'))
print('value', d['value'], 'e2e', d['e2e']['value'])
print(json.dumps(d.get('inter_video'), indent=1))
print('streaming', d['streaming']['value'], d['streaming']['e2e']['value'])
PY
