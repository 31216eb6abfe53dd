#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:icp_sweep_coop -s 0 -c 6 -o gpurun_out/r2_sweep_v2 python scripts/r2_sweep_probe.py "JOIN=4" > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
