#!/bin/bash
# round 2, session 3j: one full greedy DBS pass (25.2 M candidates), launch list of the device-timed region,
# warm-cache ncu captures of k_eval / k_commit
set -u
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --dbs-full > gpurun_out/r3_bench_dbsfull.json 2> gpurun_out/r3_bench_dbsfull.err
echo "dbs-full rc=$?"; tail -3 gpurun_out/r3_bench_dbsfull.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3_bench_dbsfull.json'))
print(json.dumps(d.get('dbs',{}).get('full_pass'))[:900]); print(json.dumps(d.get('parity_check')))
PY
# launch list: skip the reset + e2e loops (lazy recon, so ~2 launches per vectorised step), catch the value region
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/r3_launches_value.csv \
   python bench.py --steps 2 --warmup 3 --rollout 128 --no-cpu-baseline --no-extras > gpurun_out/r3_ncu_launches_value.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none --cache-control none -k regex:"k_eval_t|k_commit_t" -s 40 -c 2 \
    -f -o gpurun_out/prof_r3_delta python scripts/prof_step.py > gpurun_out/ncu_r3_delta.log 2>&1
echo "ncu delta rc=$?"
if [ -f gpurun_out/prof_r3_delta.ncu-rep ]; then
  ncu -i gpurun_out/prof_r3_delta.ncu-rep --page raw --csv > gpurun_out/prof_r3_delta_raw.csv 2>/dev/null
  rm -f gpurun_out/prof_r3_delta.ncu-rep
fi
