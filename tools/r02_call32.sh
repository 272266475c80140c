#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_fold_norm_gpu.py -m gpu -x -q -k "stream or fold or pair" 2>&1 | tail -3
T1_BENCH_VARIANTS='[{}, {}, {"decode_pair": true}]' timeout 240 python tools/decode_bench.py 2 200 2>&1 | tail -3 | cut -c1-220 | tee gpurun_out/r02_decode_producer_nodiv.txt
