#!/bin/bash
# secondary configs (BASELINE configs 0, 2, 3, 4): python scripts/bench_configs.py -> profiles/<round>_configs.json
# (compute-sanitizer is closed on this GPU pool; memory safety is covered by the bit-identity and golden tests)
set -u
mkdir -p gpurun_out
timeout 90 python scripts/bench_configs.py > gpurun_out/configs.json 2> gpurun_out/configs.err
echo "configs rc=$?"; cut -c1-300 gpurun_out/configs.json; tail -2 gpurun_out/configs.err
