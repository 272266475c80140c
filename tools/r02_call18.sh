#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attn or attention or fa_" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_engine_gpu.py tests/test_parity_7b_width_gpu.py tests/test_fullsize_gpu.py tests/test_multigroup_gpu.py -x -q 2>&1 | tail -3
timeout 500 python bench.py --steps 5 --warmup 3 --no-gpu-reference --no-cpu-baseline --phase-times --kernel-times > gpurun_out/r02_bench_attn_order.log 2> gpurun_out/r02_bench_attn_order.err
tail -1 gpurun_out/r02_bench_attn_order.log | cut -c1-200; grep -E "phase ms" gpurun_out/r02_bench_attn_order.err; grep "kernel ms" gpurun_out/r02_bench_attn_order.err | cut -c1-330
