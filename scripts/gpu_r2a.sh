#!/bin/bash
# round 2, first GPU session: correctness of the rewritten delta / observation kernels, error
# distributions against the float64 oracle, timing of the variants.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -12 gpurun_out/pytest_gpu.log
timeout 600 python scripts/parity_error_dist.py gpurun_out/parity_error_dist.json 2>&1 | tail -12
timeout 600 python scripts/r2_tune.py ${TUNE_ARGS:-} 2>&1 | tail -20 | tee gpurun_out/r2_tune.log
timeout 300 python scripts/r2_obs.py 2>&1 | tail -3 | tee gpurun_out/r2_obs.log
