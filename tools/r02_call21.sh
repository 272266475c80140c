#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_adamw_coresident.txt
: > $OUT
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -x -q -k "adamw or optimizer" 2>&1 | tail -2 | tee -a $OUT
for c in 1 2 4; do T1_ADAMW_CTAS_PER_SM=$c timeout 120 python tools/adamw_bench.py 2>&1 | tail -1 | tee -a $OUT; done
for c in 1 4 1 4; do
  echo "## bench T1_ADAMW_CTAS_PER_SM=$c" | tee -a $OUT
  T1_ADAMW_CTAS_PER_SM=$c timeout 400 python bench.py --steps 4 --warmup 3 --no-gpu-reference --no-cpu-baseline --phase-times 2>> $OUT | tail -1 | cut -c1-170 | tee -a $OUT
done
grep "phase ms" $OUT
