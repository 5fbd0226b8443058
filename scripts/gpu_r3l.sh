#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/r3_dbs_k.py 2>&1 | head -1 | tee gpurun_out/r3_dbs_k2.log
timeout 1200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3_bench_dbs3.json 2> gpurun_out/r3_bench_dbs3.err
echo "bench rc=$?"; tail -3 gpurun_out/r3_bench_dbs3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3_bench_dbs3.json'))
print(json.dumps(d.get('dbs'))[:1500])
PY
timeout 900 python -m pytest tests -m gpu -q -x -k "dbs" 2>&1 | tail -4
