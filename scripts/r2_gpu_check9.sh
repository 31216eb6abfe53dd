#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest gicp + fullsize gicp + shims + compat"
timeout 2400 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_shims.py "tests/test_gpu_fullsize.py::test_config1_full_size_gicp_vs_oracle" "tests/test_gpu_fullsize.py::test_config3_33_consecutive_64k_sweeps_through_replay" "tests/test_gpu_parity.py::test_pcl_compat_octree_map_mode_matches_the_oracle" -m gpu -q --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
echo "=== gicp batch probe"
B2ICP_GICP_DEBUG=1 timeout 900 python scripts/gicp_batch_probe.py > gpurun_out/gicp_batch.json 2> gpurun_out/gicp_batch.err
echo "probe rc=$?"; cat gpurun_out/gicp_batch.json; grep "GICP batch" gpurun_out/gicp_batch.err | tail -6
