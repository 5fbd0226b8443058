#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/r4l_bench.json 2> gpurun_out/r4l_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r4l_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r4l_bench.json'))
for k in ('value','ms_per_step','gpu_launches','roofline','roofline_eval','roofline_step_chain','e2e','dbs','clocks'):
    print(k, json.dumps(d.get(k))[:700])
PY
timeout 300 python scripts/prof_rollout.py > gpurun_out/r4l_rollout_plain.log 2>&1 && cat gpurun_out/r4l_rollout_plain.log &&
timeout 900 ncu --set full --import-source on --clock-control none --cache-control none -k regex:"k_rollout_t" -s 2 -c 1 \
    -f -o gpurun_out/prof_r4_rollout python scripts/prof_rollout.py > gpurun_out/r4l_ncu_rollout.log 2>&1
echo "ncu rollout rc=$?"; tail -3 gpurun_out/r4l_ncu_rollout.log
if [ -f gpurun_out/prof_r4_rollout.ncu-rep ]; then
  ncu -i gpurun_out/prof_r4_rollout.ncu-rep --page raw --csv > gpurun_out/prof_r4_rollout_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_r4_rollout.ncu-rep --page source --csv > gpurun_out/prof_r4_rollout_source.csv 2>/dev/null
  rm -f gpurun_out/prof_r4_rollout.ncu-rep
fi
