#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_kernels_gpu.py -x -q -k "block_local" 2>&1 | tail -5
timeout 200 python tools/vit_attn_bench.py 2>&1 | tee gpurun_out/r02_vit_attn_bench.txt
timeout 200 python -m pytest tests/test_engine_gpu.py tests/test_parity_7b_width_gpu.py -x -q 2>&1 | tail -3
