#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "norm" 2>&1 | tail -3
timeout 200 python tools/norm_bench.py 2>&1 | tee gpurun_out/r02_norm_bench.txt
timeout 400 python -m pytest tests/test_engine_gpu.py tests/test_parity_7b_width_gpu.py tests/test_multigroup_gpu.py tests/test_fold_norm_gpu.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 4 --warmup 3 --no-gpu-reference --no-cpu-baseline --phase-times --kernel-times > gpurun_out/r02_bench_norm.log 2> gpurun_out/r02_bench_norm.err
tail -1 gpurun_out/r02_bench_norm.log | cut -c1-200; grep -E "phase ms" gpurun_out/r02_bench_norm.err; grep "kernel ms" gpurun_out/r02_bench_norm.err | cut -c1-400
