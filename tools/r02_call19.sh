#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_gemm2_bn.txt
: > $OUT
timeout 240 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -3 | tee -a $OUT
T1_GEMM2_BN=192 timeout 240 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -3 | tee -a $OUT
T1_GEMM2_BN=128 timeout 240 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -3 | tee -a $OUT
for mode in auto 256; do
  echo "## T1_GEMM2_BN=$mode" | tee -a $OUT
  if [ $mode = auto ]; then timeout 200 python tools/kernel_bench.py gemm 2>&1 | tee -a $OUT | cut -c1-170; else T1_GEMM2_BN=$mode timeout 200 python tools/kernel_bench.py gemm 2>&1 | tee -a $OUT | cut -c1-170; fi
done
for mode in auto 256 auto 256; do
  echo "## bench T1_GEMM2_BN=$mode" | tee -a $OUT
  if [ $mode = auto ]; then timeout 400 python bench.py --steps 4 --warmup 3 --no-gpu-reference --no-cpu-baseline 2>> $OUT | tail -1 | cut -c1-170 | tee -a $OUT; else T1_GEMM2_BN=$mode timeout 400 python bench.py --steps 4 --warmup 3 --no-gpu-reference --no-cpu-baseline 2>> $OUT | tail -1 | cut -c1-170 | tee -a $OUT; fi
done
