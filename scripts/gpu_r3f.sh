#!/bin/bash
# round 2, session 3f: full GPU test suite, full bench (cpu baseline + comparators), reference arm, launch list
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python bench.py > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r3_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r3_bench_reference.json 2> gpurun_out/r3_bench_reference.err
echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3_bench.json'))
for k in ('value','ms_per_step','roofline','roofline_commit','roofline_step','roofline_propagate','e2e','parity_check','dbs','extra','clocks'):
    print(k, json.dumps(d.get(k))[:700])
print('cpu', json.dumps(d.get('cpu_baseline'))[:1200])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3_launches.csv \
   python bench.py --steps 1 --warmup 3 --rollout 64 --no-cpu-baseline --no-extras > gpurun_out/r3_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
