#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/r2_prop.py 2>&1 | tee gpurun_out/r2_prop2.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err
echo "bench rc=$?"; tail -5 gpurun_out/bench_dbg.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_dbg.json'))
    for k in ('value','ms_per_step','roofline','roofline_commit','roofline_step','roofline_propagate','e2e','parity_check','dbs','group','sharded_sweep','extra','clocks'):
        print(k, json.dumps(d.get(k))[:900])
    print('cpu', json.dumps(d.get('cpu_baseline'))[:900])
except Exception as e:
    print('no json', e)
PY
