python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 16 --warmup 3 > gpurun_out/bench_r01_final.json 2> gpurun_out/b.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r01_final.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'sync',round(d['config']['synchronous_call_scans_per_s']),'frac',round(d['roofline']['frac'],4),'nn',round(d['nn_search']['ms'],3),'cpu',round(d['cpu_baseline']['value'],1),'launches',d['gpu_launches'])"
ncu --set full --clock-control none --import-source on -k regex:icp_sweep_p2p -s 420 -c 8 -o gpurun_out/r01_sweep_streamed python bench.py --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/ncu_s10.log 2>&1
tail -2 gpurun_out/ncu_s10.log | cut -c1-150
