#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_gicp.py "tests/test_gpu_fullsize.py::test_config1_full_size_gicp_vs_oracle" -m gpu -q --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
B2ICP_GICP_DEBUG=1 timeout 900 python scripts/gicp_batch_probe.py > gpurun_out/gicp_batch.json 2> gpurun_out/gicp_batch.err
echo "probe rc=$?"; cat gpurun_out/gicp_batch.json; grep "GICP batch" gpurun_out/gicp_batch.err | tail -5
