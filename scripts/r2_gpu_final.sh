#!/bin/bash
# full GPU suite + smoke + N=1 bench (+ latency stream)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python scripts/latency_stream.py 2>&1 | tail -1 | tee gpurun_out/latency.json | cut -c1-300
bash scripts/r2_gpu_bench.sh 1
python -c "
import json; d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1]); r=d['roofline']
print('nn_search', r['nn_search']); print('gicp', r.get('gicp')); print('host', r['details'].get('host_ms_per_step'))"
