#!/bin/bash
# multi-GPU bench lines (N given as $1): repo arm under torchrun, then the reference arm
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r4_bench_${N}gpu.json 2> gpurun_out/r4_bench_${N}gpu.err
echo "bench rc=$?"; tail -3 gpurun_out/r4_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r4_bench_${N}gpu.json').read().strip().splitlines()[-1])
for k in ('n_gpus','value','ms_per_step','e2e','group','sharded_sweep','dbs','extra','clocks'):
    print(k, json.dumps(d.get(k))[:700])
PY
