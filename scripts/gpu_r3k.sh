#!/bin/bash
# round 2, session 3k: greedy DBS with cross-colour-group multi-accept windows: parity tests + dbs block of the bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -5 gpurun_out/pytest_gpu.log
timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3_bench_dbs2.json 2> gpurun_out/r3_bench_dbs2.err
echo "bench rc=$?"; tail -3 gpurun_out/r3_bench_dbs2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3_bench_dbs2.json'))
print(json.dumps(d.get('dbs'))[:1500]); print('value', d.get('value'))
PY
