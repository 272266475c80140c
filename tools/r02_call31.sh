#!/bin/bash
# paired decode GEMMs (t1_stream_gemm_fold_pair): parity, then the decode step pair on / off
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_fold_norm_gpu.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02_pair_tests.log
if grep -q "failed\|Timeout\|Terminated\|error" gpurun_out/r02_pair_tests.log; then echo "TESTS NOT GREEN: skipping the bench"; exit 1; fi
T1_BENCH_VARIANTS='[{"decode_pair": false}, {"decode_pair": true}, {"decode_pair": false}, {"decode_pair": true}]' timeout 240 python tools/decode_bench.py 2 200 2>&1 | tail -4 | cut -c1-220 | tee gpurun_out/r02_decode_pair_ab.txt
