#!/bin/bash
# source-level ncu capture of the batched covariance kernel (second knn_cov_kernel launch = the 32 source clouds)
set -u
mkdir -p gpurun_out
cat > /tmp/gicp_one.py <<'PY'
import sys
sys.path.insert(0, '.')
import bench
from icpslam_b200 import registration as R
map_xyzw, sweeps = bench.load_workload(0, 32)
reg = R.Registration(preset=R.PRESET_MAPPER, mode=R.MODE_GICP_BFGS)
reg.setInputTarget(map_xyzw)
rc, res = reg.alignBatch(sweeps[:32])
print(rc)
PY
timeout 900 ncu -k regex:knn_cov_kernel --launch-skip 1 -c 1 --set full --import-source on --clock-control none -o gpurun_out/r02_knn_cov -f python /tmp/gicp_one.py > gpurun_out/knn_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/knn_ncu.log
