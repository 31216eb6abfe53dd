#!/bin/bash
set -u
mkdir -p gpurun_out
GROUPS_LIST=4 bash scripts/r2_gpu_check19.sh 2>&1 | grep -v "knn\|cov_svd\|nn_cov" | head -12
bash scripts/r2_gpu_check22.sh
