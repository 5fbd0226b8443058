#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python scripts/r2_prop.py run 1024 | cut -c1-330
python scripts/r2_prop.py run 896 | cut -c1-330
