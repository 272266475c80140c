#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -4
timeout 300 python -m pytest tests/test_engine_gpu.py tests/test_parity_7b_width_gpu.py -x -q 2>&1 | tail -4
bash tools/gemm_epi_ab.sh
