#!/bin/bash
# round 2, session 4n: final evidence run with the persistent rollout kernel on the value path
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r4n_pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/r4n_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4n_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r4n_smoke.log
timeout 1500 python bench.py > gpurun_out/r4n_bench.json 2> gpurun_out/r4n_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r4n_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r4n_bench_reference.json 2> gpurun_out/r4n_bench_reference.err
echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r4n_bench.json'))
for k in ('value','ms_per_step','gpu_launches','roofline','roofline_eval','roofline_commit','roofline_step_chain','roofline_propagate','e2e','parity_check','dbs','clocks'):
    print(k, json.dumps(d.get(k))[:600])
PY
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r4n_launches.csv \
   python bench.py --steps 2 --warmup 3 --rollout 64 --no-cpu-baseline --no-extras > gpurun_out/r4n_ncu_launches.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/r4n_launches.csv
