#!/bin/bash
# round 2, third GPU session: FFT pass variants + ncu of the new passes
set -u
mkdir -p gpurun_out
timeout 900 python scripts/r2_prop.py 2>&1 | tee gpurun_out/r2_prop.log
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -5 gpurun_out/pytest_gpu.log
timeout 120 python scripts/prof_prop.py > gpurun_out/prop_plain.log 2>&1 &&
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k2_" -s 6 -c 3 \
    -f -o gpurun_out/prof_fft2 python scripts/prof_prop.py > gpurun_out/ncu_fft2.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/prop_plain.log; tail -2 gpurun_out/ncu_fft2.log
ls -la gpurun_out/*.ncu-rep
if [ -f gpurun_out/prof_fft2.ncu-rep ]; then
  ncu -i gpurun_out/prof_fft2.ncu-rep --page raw --csv > gpurun_out/prof_fft2_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_fft2.ncu-rep --page source --csv > gpurun_out/prof_fft2_source.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/prof_fft2.ncu-rep); if [ "$sz" -gt 40000000 ]; then rm -f gpurun_out/prof_fft2.ncu-rep; fi
fi
