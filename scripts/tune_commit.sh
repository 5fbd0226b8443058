#!/bin/bash
for v in 0 1 2 3 4; do
  BHOLO_COMMIT_VARIANT=$v python bench.py --steps 6 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('commit variant $v value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))"
done
