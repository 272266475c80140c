#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_gemm_2cta.txt
: > $OUT
timeout 240 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -8 | tee -a $OUT
for mode in 1 0; do
  echo "## T1_GEMM_2CTA=$mode" | tee -a $OUT
  T1_GEMM_2CTA=$mode timeout 200 python tools/kernel_bench.py gemm 2>&1 | tee -a $OUT | cut -c1-200
done
