#!/bin/bash
# A/B of the T1_SOFTMAX_POLY_EXP2 switch on a GPU box (one gpurun call, ~3 min): default build -> numbers, rebuild with the
# switch -> parity tests + numbers, rebuild default.  Usage: gpurun --timeout 600 -- 'bash tools/ab_softmax_poly_exp2.sh'
mkdir -p gpurun_out
OUT=gpurun_out/ab_softmax_poly_exp2.log
: > $OUT
run() {
  echo "## $1" | tee -a $OUT
  T1_BENCH_MAX_CTAS=148 timeout 200 python tools/decode_bench.py 2 200 2>&1 | tail -1 | tee -a $OUT
  timeout 200 python tools/decode_attn_dbg.py 2>&1 | grep '"kernel": "cluster", "S": 8' | tee -a $OUT
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --phase-times 2>> $OUT | tail -1 | cut -c1-200 | tee -a $OUT
}
run "default (MUFU.EX2 for every score)"
T1_EXTRA_NVCC=-DT1_SOFTMAX_POLY_EXP2 python -m time_r1_b200.build --force > /dev/null 2>&1 || { echo "rebuild failed" | tee -a $OUT; exit 1; }
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_multigroup_gpu.py tests/test_fullsize_gpu.py -x -q 2>&1 | tail -3 | tee -a $OUT
run "T1_SOFTMAX_POLY_EXP2 (every second exponential on the FMA pipe)"
python -m time_r1_b200.build --force > /dev/null 2>&1
echo "default build restored" | tee -a $OUT
