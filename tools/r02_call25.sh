#!/bin/bash
# L2 prefetch by a dedicated warp (off the critical path): per-line (mode 1) vs bulk (mode 2), issued by the qkv or the o GEMM
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "stream" 2>&1 | tail -3
T1_BENCH_VARIANTS='[{}, {"decode_prefetch_mb": 32}, {"decode_prefetch_mb": 64}, {"decode_prefetch_mb": 96}, {"decode_prefetch_mb": 64, "decode_prefetch_at": "qkv"}, {"decode_prefetch_mb": 96, "decode_prefetch_at": "qkv"}, {"decode_prefetch_mb": 128, "decode_prefetch_at": "qkv"}, {"env": {"T1_STREAM_PF_MODE": 1}, "decode_prefetch_mb": 32}, {"env": {"T1_STREAM_PF_MODE": 1}, "decode_prefetch_mb": 64}, {"env": {"T1_STREAM_PF_MODE": 1}, "decode_prefetch_mb": 96, "decode_prefetch_at": "qkv"}, {}]' \
  timeout 600 python tools/decode_bench.py 2 200 2>&1 | tail -12 | cut -c1-330 | tee gpurun_out/r02_decode_l2_prefetch_sweep_v2.txt
