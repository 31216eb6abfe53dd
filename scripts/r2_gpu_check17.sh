#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_fullsize.py -m gpu -q -x -k "gicp or GICP" --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
for g in ${GROUPS_LIST:-1 2 4}; do
  echo "=== groups $g"
  B2ICP_GICP_GROUPS=$g B2ICP_GICP_DEBUG=1 timeout 600 python scripts/gicp_batch_probe.py 2> gpurun_out/gicp_g$g.err | tee gpurun_out/gicp_g$g.json | cut -c1-600
  grep "GICP batch of 32" gpurun_out/gicp_g$g.err | sed -n 2,3p
done
