#!/bin/bash
# round 2, session 4a: full GPU test suite, smoke, full bench + reference arm on the committed HEAD
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r4_pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/r4_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r4_smoke.log
timeout 1500 python bench.py > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r4_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r4_bench_reference.json 2> gpurun_out/r4_bench_reference.err
echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r4_bench.json'))
for k in ('value','ms_per_step','roofline','roofline_commit','roofline_step','roofline_propagate','e2e','parity_check','dbs','clocks'):
    print(k, json.dumps(d.get(k))[:500])
r=json.load(open('gpurun_out/r4_bench_reference.json'))
print('ref', r.get('value'), r.get('cpu_baseline'))
PY
