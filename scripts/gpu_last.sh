#!/bin/bash
# GPU tests + secondary configs (BASELINE configs 0, 2, 3, 4): scripts/bench_configs.py -> profiles/<round>_configs.json
# (compute-sanitizer is closed on this GPU pool; memory safety is covered by the bit-identity and golden tests)
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -4 gpurun_out/pytest_gpu.log
timeout 90 python scripts/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err
echo "configs rc=$?"; grep -n "config2" gpurun_out/configs.json; tail -2 gpurun_out/configs.err
