#!/bin/bash
# bench.py's streamed legs for every build under icpslam_b200/lib/variants (kernel-shape A/B; dev tool)
set -u
mkdir -p gpurun_out
for lib in icpslam_b200/lib/variants/*.so; do
  B2ICP_LIB=$PWD/$lib timeout 600 python bench.py --steps 20 --warmup 5 --no-gicp --no-pairs --cpu-sample 0 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); r=d['roofline']
print('RESULT $lib value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3),'sync',round(r['details']['synchronous_call_scans_per_s']),'avg_us',round(r['avg_launch_us'],1))"
done
