#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shims.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
python scripts/r2_pairs_prof.py 2>/dev/null | grep 'rep": [23]' | cut -c1-250
