#!/bin/bash
# bench streamed leg for every build under icpslam_b200/lib/variants
set -u
mkdir -p gpurun_out
run() {
  echo "=== $*"
  env "$@" timeout 600 python bench.py --steps 12 --warmup 4 --no-gicp --no-pairs --cpu-sample 0 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3),'sync',round(r['details']['synchronous_call_scans_per_s']),'frac',round(r['frac'],4),'avg_us',round(r['avg_launch_us'],1),'clk',d['clocks']['sm_mhz'])"
}
for lib in icpslam_b200/lib/variants/*.so; do run B2ICP_LIB=$PWD/$lib; done
