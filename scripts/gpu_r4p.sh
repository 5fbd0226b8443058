#!/bin/bash
# look-ahead variant of the rollout kernel and the hybrid single-image DBS (both off by default until verified here)
set -u
mkdir -p gpurun_out
export BHOLO_ROLLOUT_LOOK_MAX_ENVS=2
timeout 600 python -m pytest tests -m gpu -q -x -k "rollout_kernel or batched_dbs" 2>&1 | tail -4
for lm in 2 0; do
  echo "look_max=$lm"
  BHOLO_ROLLOUT_LOOK_MAX_ENVS=$lm timeout 300 python scripts/r4_rollout_tune.py run | tee -a gpurun_out/r4_rollout_look.log
done
export BHOLO_DBS_ROLLOUT_MIN_ACCEPT=0.10
timeout 900 python -m pytest tests -m gpu -q -x -k "dbs" 2>&1 | tail -4
for thr in 0.10 2.0; do
  echo "dbs rollout threshold $thr"
  BHOLO_DBS_ROLLOUT_MIN_ACCEPT=$thr timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r4p_bench_$thr.json 2> gpurun_out/r4p_bench_$thr.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r4p_bench_$thr.json'))
print(json.dumps(d.get('dbs'))[:1200])
PY
done
