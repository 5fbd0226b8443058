#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python bench.py --dbs-full --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r4r_bench_dbsfull.json 2> gpurun_out/r4r_bench_dbsfull.err
echo "rc=$?"; tail -2 gpurun_out/r4r_bench_dbsfull.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r4r_bench_dbsfull.json'))
print(json.dumps(d['dbs'].get('full_pass'))[:900])
PY
