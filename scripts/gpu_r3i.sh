#!/bin/bash
# round 2, session 3i (2 GPUs): the bench under torchrun -- group / sharded-sweep blocks, NCCL gathers -- and the
# reference arm launched the same way
set -u
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r3_bench_2gpu.json 2> gpurun_out/r3_bench_2gpu.err
echo "bench2 rc=$?"; tail -5 gpurun_out/r3_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/r3_bench_ref_2gpu.json 2> gpurun_out/r3_bench_ref_2gpu.err
echo "ref2 rc=$?"; cat gpurun_out/r3_bench_ref_2gpu.json | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3_bench_2gpu.json'))
for k in ('value','n_gpus','ms_per_step','e2e','dbs','group','sharded_sweep','extra'):
    print(k, json.dumps(d.get(k))[:900])
PY
