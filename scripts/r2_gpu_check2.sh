#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "=== knob sweep"
B2ICP_DUMP_ITERS=1 timeout 900 python scripts/r2_sweep_probe.py "" "JOIN=0" "JOIN=8" "W=8" "W=8,JOIN=0" "SORT=0" "QPT=16" "QPT=1" "PROBE=0.4" "PROBE=1.5" > gpurun_out/knobs.jsonl 2> gpurun_out/knobs.err
echo "knobs rc=$?"; cat gpurun_out/knobs.jsonl; grep "per iteration" gpurun_out/knobs.err | awk 'NR%3==0' | cut -c1-400
echo "=== ncu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:icp_sweep_coop -s 4 -c 3 -o gpurun_out/r2_sweep_v1 python scripts/r2_sweep_probe.py "" > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
