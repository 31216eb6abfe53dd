#!/bin/bash
# Round-2 profile captures (one GPU): launch list of the bench command, full captures of the sweep and of the stand-alone search
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-pairs --no-gicp --cpu-sample 0"
$CMD > /dev/null 2>&1   # workload cache
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"; python scripts/ncu_summary.py list gpurun_out/r02_launches.csv | head -14
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nn_search_coop -c 2 -o gpurun_out/r02_nn_search_coop $CMD > gpurun_out/ncu_nn.log 2>&1
echo "nn rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:icp_sweep_p2p -s 100 -c 8 -o gpurun_out/r02_sweep $CMD > gpurun_out/ncu_sweep.log 2>&1
echo "sweep rc=$?"
ls -la gpurun_out/*.ncu-rep
