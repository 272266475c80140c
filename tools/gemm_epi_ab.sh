#!/bin/bash
# A/B of the GEMM epilogue (TMA tensor store vs direct register stores): microbench + one bench step each
mkdir -p gpurun_out
OUT=gpurun_out/r02_gemm_epi_ab.txt
: > $OUT
for mode in tma direct; do
  echo "## T1_GEMM_EPI=$mode" | tee -a $OUT
  T1_GEMM_EPI=$mode timeout 300 python tools/kernel_bench.py gemm 2>&1 | tee -a $OUT | cut -c1-200
  T1_GEMM_EPI=$mode timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference --phase-times 2>> $OUT | tail -1 | tee -a $OUT | cut -c1-200
done
grep "phase ms" $OUT
