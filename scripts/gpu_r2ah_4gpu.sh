set -x
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2ah_bench_4gpu.json 2> gpurun_out/r2ah_bench_4gpu.err; echo "bench4 rc=$?"
tail -3 gpurun_out/r2ah_bench_4gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2ah_bench_4gpu.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'streaming', d['streaming']['value'])
iv=d['inter_video']; print('inter', iv['value'], 'intra same batch', iv['intra_same_batch']['value'], 'loss', iv['loss_vs_intra_same_batch'], 'allgather iso ms', iv['all_gather']['ms_isolated'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 4 --steps 1 --warmup 0 > gpurun_out/r2ah_bench_ref_4gpu.json 2> gpurun_out/r2ah_bench_ref_4gpu.err; echo "ref4 rc=$?"; cut -c1-200 gpurun_out/r2ah_bench_ref_4gpu.json
