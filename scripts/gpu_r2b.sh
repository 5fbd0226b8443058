#!/bin/bash
# round 2, second GPU session: register-resident FFT passes (correctness + timing), tightened parity tests
set -u
mkdir -p gpurun_out
timeout 600 python scripts/r2_prop.py 2>&1 | tee gpurun_out/r2_prop.log
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -25 gpurun_out/pytest_gpu.log
timeout 300 python scripts/r2_tune.py 2>&1 | tail -4 | tee gpurun_out/r2_tune.log
