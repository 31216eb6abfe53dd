#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== memcheck smoke (tiny path)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck.log
echo "=== pytest gpu (all)"
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 --durations=5 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -14 gpurun_out/pytest_gpu.log | cut -c1-250
echo "=== latency"
python scripts/latency_stream.py 2>&1 | tail -1
B2ICP_NO_TINY=1 python scripts/latency_stream.py 2>&1 | tail -1
