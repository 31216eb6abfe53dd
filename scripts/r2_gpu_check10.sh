#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shims.py::test_pointcloud2_payload_to_cloud -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_gicp_launches.csv python scripts/r2_gicp_prof.py 32 > gpurun_out/ncu_gicp.log 2>&1
python scripts/ncu_summary.py list gpurun_out/r2_gicp_launches.csv | head -16
