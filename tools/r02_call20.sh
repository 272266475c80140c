#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_gemm2_bn_v2.txt
: > $OUT
for mode in auto 256 auto 256; do
  echo "## bench T1_GEMM2_BN=$mode (narrow tiles only for M <= 2048 or K <= 2048)" | tee -a $OUT
  if [ $mode = auto ]; then timeout 400 python bench.py --steps 4 --warmup 3 --no-gpu-reference --no-cpu-baseline 2>> $OUT | tail -1 | cut -c1-170 | tee -a $OUT; else T1_GEMM2_BN=$mode timeout 400 python bench.py --steps 4 --warmup 3 --no-gpu-reference --no-cpu-baseline 2>> $OUT | tail -1 | cut -c1-170 | tee -a $OUT; fi
done
