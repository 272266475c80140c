#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_fold_norm_gpu.py -x -q > gpurun_out/norm_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/norm_tests.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r01_launches_bench_v7.csv python bench.py --ncu-range --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r01_launches_bench_v7.csv
timeout 200 python bench.py --steps 5 --warmup 3 --phase-times --kernel-times > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-250; grep "phase ms" gpurun_out/bench.err
