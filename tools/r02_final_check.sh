#!/bin/bash
# what the driver runs at round end: GPU test suite, smoke(), the default bench and the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_gpu_tests_final.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee -a gpurun_out/r02_gpu_tests_final.log
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_final.log 2> gpurun_out/r02_bench_final.err ) 2>&1 | grep real
tail -1 gpurun_out/r02_bench_final.log | cut -c1-300
