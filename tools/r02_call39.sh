#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "decode_attention" 2>&1 | tail -2
T1_BENCH_VARIANTS='[{}, {}]' timeout 200 python tools/decode_bench.py 2 200 2>&1 | tail -2 | cut -c1-200 | tee gpurun_out/r02_decode_wide_merge_batched.txt
