#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 600 python scripts/r2_sweep_probe.py "" "NO_STAGE=1" "QPT=4" "QPT=16" "QPT=16,NO_STAGE=1" 2> gpurun_out/knobs.err | tee gpurun_out/knobs.jsonl
grep "per iteration" gpurun_out/knobs.err | cut -c1-700
