#!/bin/bash
set -u
mkdir -p gpurun_out
for ms in 100 200 1000 100 200 1000; do
  B2_BENCH_CLOCK_MS=$ms timeout 600 python bench.py --steps 20 --warmup 5 --no-gicp --no-pairs --cpu-sample 0 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); r=d['roofline']
print('RESULT clock_ms $ms value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3),'host',r['details'].get('host_ms_per_step'),'samples',d['clocks']['samples'])"
done
