#!/bin/bash
# final check of the committed defaults: GPU tests, smoke, full bench, reference arm
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r4q_pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/r4q_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4q_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r4q_smoke.log
timeout 1500 python bench.py > gpurun_out/r4q_bench.json 2> gpurun_out/r4q_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r4q_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r4q_bench_reference.json 2> gpurun_out/r4q_bench_reference.err
echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r4q_bench.json'))
for k in ('value','ms_per_step','gpu_launches','roofline','roofline_eval','roofline_commit','roofline_step_chain','roofline_propagate','e2e','parity_check','dbs','sharded_sweep','group','clocks'):
    print(k, json.dumps(d.get(k))[:500])
c=d['cpu_baseline']; print('cpu', c['value'], c.get('tuned_allocator'))
PY
