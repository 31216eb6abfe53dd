#!/bin/bash
set -u
mkdir -p gpurun_out
B2ICP_GICP_DEBUG=1 timeout 600 python scripts/r2_gicp_prof.py 8 2>&1 | tail -12
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_gicp_launches.csv python scripts/r2_gicp_prof.py 8 > gpurun_out/ncu_gicp.log 2>&1
python scripts/ncu_summary.py list gpurun_out/r2_gicp_launches.csv | head -24
