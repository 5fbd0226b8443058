#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python scripts/r2_tune.py run | tee gpurun_out/r4_commit_tune.log
BHOLO_LIB=build/exp/lib_prev2.so python scripts/r2_tune.py run | tee -a gpurun_out/r4_commit_tune.log
timeout 300 python scripts/r3_dbs_k.py 2>&1 | tee gpurun_out/r4_dbs_k.log
echo prev
BHOLO_LIB=build/exp/lib_prev2.so timeout 300 python scripts/r3_dbs_k.py 2>&1 | head -6 | tee gpurun_out/r4_dbs_k_prev.log
