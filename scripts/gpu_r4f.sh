#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python scripts/r4_l2pf.py 2>&1 | tee gpurun_out/r4_l2pf2.log
