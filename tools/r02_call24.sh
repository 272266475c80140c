#!/bin/bash
# L2 prefetch of the gate_up weights during the latency-bound part of the decode layer: parity of the decode path + sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -x -q -k "stream or decode or rollout" 2>&1 | tail -5 | tee gpurun_out/r02_pf_tests.log
T1_BENCH_VARIANTS='[{}, {"decode_prefetch_mb": 32}, {"decode_prefetch_mb": 64}, {"decode_prefetch_mb": 96}, {"decode_prefetch_mb": 128}, {"decode_prefetch_mb": 64, "decode_prefetch_at": "qkv"}, {"decode_prefetch_mb": 96, "decode_prefetch_at": "qkv"}, {"decode_prefetch_mb": 64, "decode_attn": "wide"}, {}]' \
  timeout 600 python tools/decode_bench.py 2 200 2>&1 | tail -12 | tee gpurun_out/r02_decode_l2_prefetch_sweep.txt
