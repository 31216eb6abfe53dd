#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gicp + fullsize"
timeout 2400 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_fullsize.py -m gpu -q --timeout 1200 --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "=== gicp batch probe"
timeout 900 python scripts/gicp_batch_probe.py > gpurun_out/gicp_batch.json 2> gpurun_out/gicp_batch.err
echo "probe rc=$?"; cat gpurun_out/gicp_batch.json; tail -3 gpurun_out/gicp_batch.err
