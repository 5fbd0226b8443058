#!/bin/bash
# GPU tests, k_eval variant timings, then the bench line without the CPU legs
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python scripts/tune_eval.py > gpurun_out/tune_eval.log 2>&1
echo "tune rc=$?"; cat gpurun_out/tune_eval.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_quick.json'))
    print('value',round(d['value']),'e2e',round(d['e2e']['value']),'roofline',round(d['roofline']['achieved']),round(d['roofline']['frac'],3),'eval_ms',d['roofline']['ms_per_launch'],'ms/step',d['ms_per_step'])
    print(d['extra']); print(d['clocks'])
except Exception as e:
    print('no bench json', e); print(open('gpurun_out/bench_quick.err').read()[-2000:])
PY
