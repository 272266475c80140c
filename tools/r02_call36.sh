#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_fold_norm_gpu.py -m gpu -x -q -k "stream or fold or pair" 2>&1 | tail -3
T1_STREAM_FAST_SYNC=1 timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_fold_norm_gpu.py -m gpu -x -q -k "stream or fold or pair" 2>&1 | tail -2
T1_BENCH_VARIANTS='[{"env": {"T1_STREAM_FAST_SYNC": 0}}, {"env": {"T1_STREAM_FAST_SYNC": 1}}, {"env": {"T1_STREAM_FAST_SYNC": 0}}, {"env": {"T1_STREAM_FAST_SYNC": 1}}, {"env": {"T1_STREAM_FAST_SYNC": 0, "T1_STREAM_EARLY_RESID": 0}}]' timeout 500 python tools/decode_bench.py 2 200 2>&1 | tail -5 | cut -c1-215 | tee gpurun_out/r02_decode_fast_sync.txt
