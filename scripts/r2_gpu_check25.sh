#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --no-pairs > gpurun_out/bench_gicp.json 2> gpurun_out/bench_gicp.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_gicp.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/bench_gicp.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value',round(d['value']),'e2e',round(d['e2e']['value'])); print('gicp', r.get('gicp'))"
bash scripts/r2_gpu_check24.sh
