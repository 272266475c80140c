#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fold_norm_gpu.py tests/test_answer_scan_gpu.py "tests/test_kernels_gpu.py::test_rmsnorm_fwd_bwd" -x -q > gpurun_out/fold_tests.log 2>&1; RC=$?
tail -5 gpurun_out/fold_tests.log
for F in 0 1; do
  T1_DECODE_FOLD=$F T1_BENCH_MAX_CTAS=148 timeout 200 python tools/decode_bench.py 2 200 2>&1 | tail -1 | sed "s/^/fold=$F /" | tee -a gpurun_out/decode_fold.log
done
timeout 330 python bench.py --phase-times --kernel-times > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-300; grep "phase ms" gpurun_out/bench.err
if [ $RC -eq 0 ]; then
  T1_DECODE_FOLD=1 timeout 330 python bench.py --phase-times > gpurun_out/bench_fold.log 2> gpurun_out/bench_fold.err
  tail -1 gpurun_out/bench_fold.log | cut -c1-300; grep "phase ms" gpurun_out/bench_fold.err
fi
