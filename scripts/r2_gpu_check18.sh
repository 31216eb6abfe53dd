#!/bin/bash
# launch list of one GICP batch: (a) the round kernels, (b) the source-side setup kernels
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
B2ICP_GICP_GROUPS=${1:-4} timeout 900 ncu -k regex:gicp_ --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 80 --csv --log-file gpurun_out/gicp_round_launches.csv python /tmp/gicp_one.py > gpurun_out/gicp_setup.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/gicp_setup.log
python scripts/ncu_list.py gpurun_out/gicp_round_launches.csv 2>/dev/null | cut -c1-110 | tail -45
B2ICP_GICP_GROUPS=${1:-4} timeout 900 ncu -k regex:knn_ --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 12 --csv --log-file gpurun_out/gicp_knn_launches.csv python /tmp/gicp_one.py > gpurun_out/gicp_setup2.log 2>&1
python scripts/ncu_list.py gpurun_out/gicp_knn_launches.csv 2>/dev/null | cut -c1-110
