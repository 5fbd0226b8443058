#!/bin/bash
# One gpurun call: GPU tests, the bench line, then the ncu launch list and one
# full capture of the dominant kernel (each only after the plain run exited 0).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; cat gpurun_out/bench_ref.json
if [ "${NCU:-1}" = "1" ]; then
  SMALL="python bench.py --steps 2 --warmup 1 --rollout 16 --no-cpu-baseline"
  timeout 600 $SMALL > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
      --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  timeout 600 $SMALL > gpurun_out/bench_small2.json 2> gpurun_out/bench_small2.err &&
  timeout 1200 ncu --set full --clock-control none --import-source on \
      -k regex:"k_(eval|commit|rows_fwd|cols|rows_inv|loss_sums)" -s 24 -c 28 \
      -f -o gpurun_out/prof_hotpath $SMALL > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
fi
