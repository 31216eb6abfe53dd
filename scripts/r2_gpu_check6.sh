#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
bash scripts/r2_gpu_bench.sh 1
B2ICP_NO_GRAPH=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-pairs --cpu-sample 0 > gpurun_out/bench_nograph.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/bench_nograph.json').read().strip().splitlines()[-1]); print('NO_GRAPH value',round(d['value']),'e2e',round(d['e2e']['value']))"
python scripts/latency_stream.py 2>&1 | tail -3
