#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/prof_step.py > gpurun_out/prof_step_plain.log 2>&1
cat gpurun_out/prof_step_plain.log
timeout 600 ncu --set full --import-source on --clock-control none --cache-control none -k regex:"k_step_fused" -s 10 -c 1 \
    -f -o gpurun_out/prof_r3_step python scripts/prof_step.py > gpurun_out/ncu_r3_step.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_r3_step.log
if [ -f gpurun_out/prof_r3_step.ncu-rep ]; then
  ncu -i gpurun_out/prof_r3_step.ncu-rep --page raw --csv > gpurun_out/prof_r3_step_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_r3_step.ncu-rep --page source --csv > gpurun_out/prof_r3_step_source.csv 2>/dev/null
  rm -f gpurun_out/prof_r3_step.ncu-rep
fi
