#!/bin/bash
# build variants locally (nvcc), run them on the GPU: bash scripts/tune_fft.sh build | run
set -u
VARIANTS="2,4,8,1 2,2,8,1 4,2,8,1 4,1,8,1 1,4,8,1 2,3,8,1 2,4,4,2 2,4,4,1 2,4,16,1"
if [ "$1" = build ]; then
  mkdir -p build
  for v in $VARIANTS; do IFS=, read rw rm cw cm <<< "$v"
    nvcc -DBH_ROWS_W=$rw -DBH_ROWS_MINB=$rm -DBH_COLS_W=$cw -DBH_COLS_MINB=$cm -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -cudart static -o build/libbholo_fft_${rw}_${rm}_${cw}_${cm}.so binary_hologram_reinforcement_learning_b200/csrc/bholo.cu &
  done; wait; ls build/libbholo_fft_*
else
  for v in $VARIANTS; do IFS=, read rw rm cw cm <<< "$v"
    echo "rows W=$rw minb=$rm cols W=$cw minb=$cm: $(BHOLO_LIB=$PWD/build/libbholo_fft_${rw}_${rm}_${cw}_${cm}.so python scripts/prof_prop.py 2>&1 | tail -2 | tr '\n' ' ')"
  done
fi
