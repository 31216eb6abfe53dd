#!/bin/bash
# N-rank diagnostics of the streamed leg: what the host does per step under different knobs
set -u
N=${1:-4}
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-pairs --cpu-sample 0 > gpurun_out/diag_$name.json 2> gpurun_out/diag_$name.err
  echo "== $name rc=$?"; grep "resident leg" gpurun_out/diag_$name.err | cut -c1-200
  python -c "
import json; d=json.loads(open('gpurun_out/diag_$name.json').read().strip().splitlines()[-1]); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3))"
}
nproc; python -c "import os; print('affinity', len(os.sched_getaffinity(0)))"; lscpu | grep -E "Model name|Thread|Core|Socket|^CPU\(s\)" 
run base A=1
run nograph B2ICP_NO_GRAPH=1
run noflush BENCH_NO_FLUSH=1
run blocking CUDA_DEVICE_SCHEDULE=blocking BENCH_NO_FLUSH=1
