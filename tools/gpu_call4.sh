#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fold_norm_gpu.py "tests/test_kernels_gpu.py" -x -q > gpurun_out/fold_tests.log 2>&1; RC=$?
tail -3 gpurun_out/fold_tests.log
for F in 0 1 0 1; do
  T1_DECODE_FOLD=$F T1_BENCH_MAX_CTAS=148 timeout 200 python tools/decode_bench.py 2 200 2>&1 | tail -1 | sed "s/^/fold=$F /" | tee -a gpurun_out/decode_fold2.log
done
