#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-streaming --no-other-workloads"
for rep in 1 2; do
timeout 300 $B > gpurun_out/r2z_bench_static_$rep.json 2>/dev/null
HVR_DEBUG_FLAGS=262144 timeout 300 $B > gpurun_out/r2z_bench_clc_$rep.json 2>/dev/null
HVR_DEBUG_FLAGS=262144 HVR_FORK_PROPOSALS=0 HVR_FORK_POST=0 timeout 300 $B > gpurun_out/r2z_bench_clc_nofork_$rep.json 2>/dev/null
HVR_FORK_PROPOSALS=0 HVR_FORK_POST=0 timeout 300 $B > gpurun_out/r2z_bench_static_nofork_$rep.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2z_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, '%.2f fps  %.3f ms  gemm %.2f ms  tensor_work_frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['tensor_work_frac']))
    except Exception as e:
        print(f, 'failed', e)
PY
