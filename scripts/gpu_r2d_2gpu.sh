set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "intervideo or ragged or stream" > gpurun_out/r2d_2gpu_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r2d_2gpu_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d_bench_2gpu.json 2> gpurun_out/r2d_bench_2gpu.err; echo "bench2 rc=$?"
cat gpurun_out/r2d_bench_2gpu.json; tail -20 gpurun_out/r2d_bench_2gpu.err
