#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/r2_prop.py 2>&1 | tee gpurun_out/r2_prop3.log
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err
echo "bench rc=$?"; tail -5 gpurun_out/bench_dbg.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_dbg.json'))
    for k in ('value','ms_per_step','roofline_propagate','dbs','group','sharded_sweep'):
        print(k, json.dumps(d.get(k))[:1500])
except Exception as e:
    print('no json', e)
PY
