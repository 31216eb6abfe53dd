#!/bin/bash
# usage: r2_gpu_bench.sh N   (N ranks on one box)
set -u
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
fi
echo "bench N=$N rc=$?"; tail -4 gpurun_out/bench_n$N.err | cut -c1-300; python -c "
import json,sys; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
r=d['roofline']
print('N',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3),'sync',round(r['details']['synchronous_call_scans_per_s']),'frac',round(r['frac'],4),'avg_us',round(r['avg_launch_us'],1),'iters',r['details']['mean_iterations'],'searched',round(r['searched_fraction'],3))
print('pairs',r.get('pairs',{}).get('pairs_per_s'), 'parity',d['parity'], 'rec_ok', r['details']['gathered_records_ok'], 'clocks', d['clocks'])"
