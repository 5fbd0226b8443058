#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r4o_bench_2gpu.json 2> gpurun_out/r4o_bench_2gpu.err
echo "bench rc=$?"; tail -3 gpurun_out/r4o_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r4o_bench_2gpu.json').read().strip().splitlines()[-1])
for k in ('n_gpus','value','ms_per_step','gpu_launches','roofline','e2e','dbs','sharded_sweep'):
    print(k, json.dumps(d.get(k))[:400])
PY
