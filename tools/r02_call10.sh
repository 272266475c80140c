#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -22 | tee gpurun_out/r02_gpu_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-gpu-reference --phase-times --kernel-times > gpurun_out/r02_bench_2cta.log 2> gpurun_out/r02_bench_2cta.err
tail -1 gpurun_out/r02_bench_2cta.log | cut -c1-900; grep -E "phase ms" gpurun_out/r02_bench_2cta.err
