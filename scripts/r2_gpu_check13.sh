#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shims.py -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
python scripts/latency_stream.py 2>&1 | tail -1 | tee gpurun_out/latency.json
bash scripts/r2_gpu_bench.sh 1
python -c "
import json; d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); r=d['roofline']
print('nn_search', r['nn_search']); print('gicp', r.get('gicp')); print('host', r['details']['host_ms_per_step'])"
