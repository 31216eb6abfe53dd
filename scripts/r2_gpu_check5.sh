#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "=== bench N=1"
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(r['details']['synchronous_call_scans_per_s']),'frac',round(r['frac'],4),'avg_us',round(r['avg_launch_us'],1),'nn_ms',r['nn_search']['ms'])
print('pairs',r.get('pairs')); print('parity',d['parity']); print('cpu',d['cpu_baseline'])"
