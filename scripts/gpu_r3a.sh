#!/bin/bash
# round 2, session 3a: packed FP32x2 FFT arithmetic vs scalar; FFMA2 throughput; parity tests; ncu of the passes
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
./build/exp/pk_bench 2>&1 | tee gpurun_out/r3_pk_bench.log
timeout 900 python scripts/r2_prop.py 2>&1 | tee gpurun_out/r3_prop.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -5 gpurun_out/pytest_gpu.log
timeout 120 python scripts/prof_prop.py > gpurun_out/prop_plain.log 2>&1 &&
timeout 900 ncu --set full --import-source on --clock-control none --cache-control none -k regex:"k2_" -s 9 -c 3 \
    -f -o gpurun_out/prof_r3_fft python scripts/prof_prop.py > gpurun_out/ncu_r3_fft.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/prop_plain.log; tail -2 gpurun_out/ncu_r3_fft.log
if [ -f gpurun_out/prof_r3_fft.ncu-rep ]; then
  ncu -i gpurun_out/prof_r3_fft.ncu-rep --page raw --csv > gpurun_out/prof_r3_fft_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_r3_fft.ncu-rep --page source --csv > gpurun_out/prof_r3_fft_source.csv 2>/dev/null
  rm -f gpurun_out/prof_r3_fft.ncu-rep
fi
