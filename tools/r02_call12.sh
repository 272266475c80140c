#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_norm_ab.txt
: > $OUT
for mode in 1 0 1 0; do
  echo "## T1_NORM_ROWS=$mode" | tee -a $OUT
  T1_NORM_ROWS=$mode timeout 200 python tools/norm_bench.py 2>&1 | tee -a $OUT
  T1_NORM_ROWS=$mode timeout 400 python bench.py --steps 4 --warmup 3 --no-gpu-reference --no-cpu-baseline --phase-times 2>> $OUT | tail -1 | cut -c1-160 | tee -a $OUT
done
grep "phase ms" $OUT
