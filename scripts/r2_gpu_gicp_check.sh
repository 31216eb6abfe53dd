#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_fullsize.py -m gpu -q -x -k "gicp or GICP or cov" --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
for g in ${GROUPS_LIST:-2 4}; do
  echo "=== groups $g"
  B2ICP_GICP_GROUPS=$g B2ICP_GICP_DEBUG=1 timeout 600 python scripts/gicp_batch_probe.py 2> gpurun_out/gicp_g$g.err > gpurun_out/gicp_g$g.json
  python -c "
import json;d=json.load(open('gpurun_out/gicp_g$g.json'));print($g,{k:round(v.get('scans_per_s',v.get('pairs_per_s'))) for k,v in d.items()})"
  grep -B4 "GICP batch of 32" gpurun_out/gicp_g$g.err | sed -n 6,10p
done
bash scripts/r2_gpu_gicp_knn_launches.sh | tail -10
