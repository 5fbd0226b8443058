#!/bin/bash
# eager observation path at N GPUs with and without binding every rank to its GPU's NUMA node
set -u
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r4_topo_${N}gpu.txt 2>&1
lscpu | grep -i "numa\|socket\|^CPU(s)" >> gpurun_out/r4_topo_${N}gpu.txt
for mode in bind nobind; do
  if [ $mode = nobind ]; then export BHOLO_NO_NUMA_BIND=1; else unset BHOLO_NO_NUMA_BIND; fi
  BHOLO_TOPO_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r4_numa_${mode}_${N}gpu.json 2> gpurun_out/r4_numa_${mode}_${N}gpu.err
  echo "$mode rc=$?"; grep "numa node" gpurun_out/r4_numa_${mode}_${N}gpu.err | head -8
  python - <<PY
import json
d=json.loads(open('gpurun_out/r4_numa_${mode}_${N}gpu.json').read().strip().splitlines()[-1])
print('$mode', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'pcie', d['e2e']['pcie_d2h_gb_per_s'], d.get('host_binding'))
PY
done
tail -12 gpurun_out/r4_topo_${N}gpu.txt
