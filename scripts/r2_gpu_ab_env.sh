#!/bin/bash
# A/B of a tuning knob on the bench's streamed leg: r2_gpu_ab_env.sh "ENV=1" ...
set -u
mkdir -p gpurun_out
for cfg in "$@"; do
  echo "=== $cfg"
  env $cfg timeout 900 python bench.py --steps 20 --warmup 5 --no-gicp --no-pairs > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); r=d['roofline']
print('RESULT value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3),'sync',round(r['details']['synchronous_call_scans_per_s']),'frac',round(r['frac'],4),'avg_us',round(r['avg_launch_us'],1),'parity',d['parity'])"
done
