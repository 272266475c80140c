#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -k layernorm -x -q 2>&1 | tail -5 | tee gpurun_out/r02_ln.log
timeout 600 python -m pytest tests/test_engine_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/r02_engine.log
timeout 900 python -m pytest tests/test_parity_7b_width_gpu.py -x -q -s 2>&1 | tail -60 | tee gpurun_out/r02_parity7b.log
timeout 600 python bench.py --verify gpurun_out/verify_n1.pt 2>&1 | tail -3 | tee gpurun_out/r02_verify_n1.log
