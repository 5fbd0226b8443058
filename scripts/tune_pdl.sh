#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for v in 0 1; do
  if [ $v = 1 ]; then export BHOLO_NO_PDL=1; fi
  python bench.py --steps 6 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('NO_PDL=$v value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'eval_ms', d['roofline']['ms_per_launch'], 'frac', round(d['roofline']['frac'],3), 'dbs', round(d['extra']['dbs_greedy_candidates_per_s']))"
done
