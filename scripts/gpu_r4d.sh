#!/bin/bash
# session 4d: full bench with the per-rank-target sweep block and the allocator-tuned CPU figures; reference arm
set -u
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r4d_bench.json 2> gpurun_out/r4d_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r4d_bench.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r4d_bench_reference.json 2> gpurun_out/r4d_bench_reference.err
echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r4d_bench.json'))
for k in ('value','ms_per_step','roofline_propagate','sharded_sweep','group'):
    print(k, json.dumps(d.get(k))[:900])
c=d['cpu_baseline']; print('cpu', c['value'], c.get('tuned_allocator'))
r=json.load(open('gpurun_out/r4d_bench_reference.json'))
print('ref', r.get('value'), r['cpu_baseline'].get('tuned_allocator'))
PY
