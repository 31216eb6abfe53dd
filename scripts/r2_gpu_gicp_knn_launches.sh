#!/bin/bash
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
timeout 900 ncu -k regex:"knn_|cov_svd" --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 12 --csv --log-file gpurun_out/gicp_knn_launches.csv python /tmp/gicp_one.py > gpurun_out/gicp_setup2.log 2>&1
python scripts/ncu_list.py gpurun_out/gicp_knn_launches.csv 2>/dev/null | cut -c1-110
