#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu (parity)"
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "=== knob sweep"
B2ICP_DUMP_ITERS=1 timeout 900 python scripts/r2_sweep_probe.py "" "TILES=0" "QPT=16" "QPT=1" "W=8" "PROBE=0.5" "PROBE=1.0" > gpurun_out/knobs.jsonl 2> gpurun_out/knobs.err
echo "knobs rc=$?"; cat gpurun_out/knobs.jsonl; grep "per iteration" gpurun_out/knobs.err | awk 'NR%3==0' | cut -c1-330
echo "=== bench"
timeout 900 python bench.py --steps 8 --warmup 3 --cpu-sample 0 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['config']['synchronous_call_scans_per_s']),'frac',round(d['roofline']['frac'],4),'avg_us',round(d['roofline']['avg_launch_us'],1),'nn_ms',d['nn_search']['ms'])"
echo "=== ncu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:icp_sweep_coop -s 0 -c 6 -o gpurun_out/r2_sweep_v3 python scripts/r2_sweep_probe.py "" > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu.log
