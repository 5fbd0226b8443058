#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python scripts/r2_prop.py 2>&1 | tee gpurun_out/r2_prop.log
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -5 gpurun_out/pytest_gpu.log
