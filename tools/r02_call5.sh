#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_engine_gpu.py -x -q -k "vit or scoring" 2>&1 | tail -4
timeout 300 python -m pytest tests/test_parity_7b_width_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/vit_attn_bench.py 2>&1 | tee gpurun_out/r02_vit_attn_bench.txt
timeout 300 python -m pytest tests/test_trainer_gpu.py tests/test_preproc_gpu.py -x -q 2>&1 | tail -4
