#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "propag or full_size or golden or simulate or resync" 2>&1 | tail -5
timeout 900 python scripts/r4_prop.py 2>&1 | tee gpurun_out/r4_prop.log
