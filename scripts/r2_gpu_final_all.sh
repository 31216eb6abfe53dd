#!/bin/bash
set -u
bash scripts/r2_gpu_final.sh
bash scripts/r2_gpu_profiles.sh
