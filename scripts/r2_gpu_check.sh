#!/bin/bash
# Round-2 GPU check: sanitizer on the smoke path, parity tests, then per-iteration device times of the bench batch.
set -u
mkdir -p gpurun_out
echo "=== memcheck smoke" 
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/memcheck.log
echo "=== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "=== bench variants"
for W in 32 8; do
  B2ICP_W=$W B2ICP_DUMP_ITERS=1 timeout 900 python bench.py --steps 4 --warmup 3 --cpu-sample 0 > gpurun_out/bench_w$W.json 2> gpurun_out/bench_w$W.err
  echo "bench W=$W rc=$?"; tail -c 600 gpurun_out/bench_w$W.json; grep "per iteration" gpurun_out/bench_w$W.err | tail -2
done
