#!/bin/bash
# One gpurun call: GPU tests, the bench line of both arms, then (NCU=1) the ncu launch list of the
# bench command and full captures of the hot kernels -- each only after the plain run exited 0.
# Every step has its own timeout; the whole script is meant to stay under 12 minutes.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1

timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log

timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/pytest_gpu.log

timeout 600 python bench.py --steps "${BENCH_STEPS:-10}" --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err

timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json

if [ "${NCU:-1}" != "1" ]; then exit 0; fi

SMALL="python bench.py --steps 2 --warmup 1 --rollout 16 --no-cpu-baseline"

# gpurun returns at most 64 MiB: keep the raw metric pages as CSV, drop the .ncu-rep files
raw_csv() {
  if [ -f "gpurun_out/$1.ncu-rep" ]; then
    ncu -i "gpurun_out/$1.ncu-rep" --page raw --csv > "gpurun_out/$1_raw.csv" 2>/dev/null
    rm -f "gpurun_out/$1.ncu-rep"
  fi
}

# (1) launch list of the bench command
timeout 300 $SMALL > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches.csv $SMALL > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"

# (2) full capture of the delta kernels (ncu default: caches flushed before every replay)
timeout 300 $SMALL > gpurun_out/bench_small2.json 2> gpurun_out/bench_small2.err &&
timeout 600 ncu --set full --clock-control none -k regex:"k_(eval_t|commit_t)" -s 20 -c 8 \
    -f -o gpurun_out/prof_hotpath $SMALL > gpurun_out/ncu_full.log 2>&1
echo "ncu delta kernels rc=$?"; tail -1 gpurun_out/ncu_full.log
raw_csv prof_hotpath

# (3) k_eval with warm caches: the live L2 residency of the impulse table
timeout 600 ncu --set full --clock-control none --cache-control none -k regex:"k_eval_t" -s 20 -c 4 \
    -f -o gpurun_out/prof_eval_warm $SMALL > gpurun_out/ncu_warm.log 2>&1
echo "ncu warm rc=$?"; tail -1 gpurun_out/ncu_warm.log
raw_csv prof_eval_warm

# (4) propagation passes
timeout 120 python scripts/prof_prop.py > gpurun_out/prop_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none \
    -k regex:"k_(rows_fwd|cols|rows_inv|loss_final)" -s 10 -c 10 \
    -f -o gpurun_out/prof_prop python scripts/prof_prop.py > gpurun_out/ncu_prop.log 2>&1
echo "ncu prop rc=$?"; cat gpurun_out/prop_plain.log
raw_csv prof_prop

# (5) bundled candidate-list evaluation (bh_eval_flips inside the bench extras)
timeout 600 ncu --set full --clock-control none --cache-control none -k regex:"k_eval_bundle" -c 2 \
    -f -o gpurun_out/prof_bundle $SMALL > gpurun_out/ncu_bundle.log 2>&1
echo "ncu bundle rc=$?"; tail -1 gpurun_out/ncu_bundle.log
raw_csv prof_bundle
