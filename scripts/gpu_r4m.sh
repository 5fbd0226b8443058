#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python scripts/r4_rollout_tune.py 2>&1 | tee gpurun_out/r4_rollout_tune.log
