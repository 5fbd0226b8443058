#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "rollout_kernel" 2>&1 | tail -15
timeout 600 python scripts/r4_rollout.py 2>&1 | tee gpurun_out/r4_rollout.log
