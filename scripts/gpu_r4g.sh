#!/bin/bash
# round 2, session 4g: final evidence run -- GPU tests, smoke, full bench + reference arm, ncu launch lists,
# ncu --set full of the step kernels and of the propagation passes (warm cache)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r4g_pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/r4g_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4g_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r4g_smoke.log
timeout 1500 python bench.py > gpurun_out/r4g_bench.json 2> gpurun_out/r4g_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r4g_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r4g_bench_reference.json 2> gpurun_out/r4g_bench_reference.err
echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r4g_bench.json'))
for k in ('value','ms_per_step','roofline','roofline_commit','roofline_step','roofline_propagate','e2e','parity_check','dbs','sharded_sweep','clocks'):
    print(k, json.dumps(d.get(k))[:600])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r4g_launches_e2e.csv \
   python bench.py --steps 1 --warmup 3 --rollout 64 --no-cpu-baseline --no-extras > gpurun_out/r4g_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/r4g_launches_value.csv \
   python bench.py --steps 2 --warmup 3 --rollout 128 --no-cpu-baseline --no-extras > gpurun_out/r4g_ncu_launches_value.log 2>&1
echo "ncu launches value rc=$?"
timeout 120 python scripts/prof_step.py > gpurun_out/r4g_step_plain.log 2>&1 &&
timeout 900 ncu --set full --import-source on --clock-control none --cache-control none -k regex:"k_eval_t|k_commit_t" -s 20 -c 4 \
    -f -o gpurun_out/prof_r4_step python scripts/prof_step.py > gpurun_out/r4g_ncu_step.log 2>&1
echo "ncu step rc=$?"; tail -2 gpurun_out/r4g_ncu_step.log
timeout 120 python scripts/prof_prop.py > gpurun_out/r4g_prop_plain.log 2>&1 &&
timeout 900 ncu --set full --import-source on --clock-control none --cache-control none -k regex:"k2_|k4_" -s 9 -c 3 \
    -f -o gpurun_out/prof_r4_fft python scripts/prof_prop.py > gpurun_out/r4g_ncu_fft.log 2>&1
echo "ncu fft rc=$?"; cat gpurun_out/r4g_prop_plain.log
for n in step fft; do
  if [ -f gpurun_out/prof_r4_$n.ncu-rep ]; then
    ncu -i gpurun_out/prof_r4_$n.ncu-rep --page raw --csv > gpurun_out/prof_r4_${n}_raw.csv 2>/dev/null
    ncu -i gpurun_out/prof_r4_$n.ncu-rep --page source --csv > gpurun_out/prof_r4_${n}_source.csv 2>/dev/null
    rm -f gpurun_out/prof_r4_$n.ncu-rep
  fi
done
