python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_warp.json 2> gpurun_out/bench_r01_warp.err
tail -1 gpurun_out/bench_r01_warp.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_v2.csv python bench.py --steps 2 --warmup 3 --cpu-sample 0 > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:icp_sweep_p2p -s 90 -c 8 -o gpurun_out/r01_sweep_warp python bench.py --steps 1 --warmup 3 --cpu-sample 0 > gpurun_out/ncu_s9.log 2>&1
ls -la gpurun_out | tail -5
