#!/bin/bash
# round 2, session 3b: TMA-pipelined step kernels (stages sweep), split-stream experiment, FFT with PDL
set -u
mkdir -p gpurun_out
timeout 900 python scripts/r3_tune.py 2>&1 | tee gpurun_out/r3_tune.log
timeout 300 python scripts/r2_prop.py 2>&1 | head -3 | tee gpurun_out/r3_prop_pdl.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -5 gpurun_out/pytest_gpu.log
