#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/r2ai_pipeline_tests.txt 2>&1; tail -3 gpurun_out/r2ai_pipeline_tests.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-streaming --no-other-workloads 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value'])"
