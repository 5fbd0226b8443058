#!/bin/bash
set -u
mkdir -p gpurun_out
./build/exp/dft_rate 2>&1 | tee gpurun_out/r4_dft_rate.log
timeout 300 python scripts/r4_sweep_profile.py 2>&1 | tee gpurun_out/r4_sweep_profile.log | head -60
