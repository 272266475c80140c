#!/bin/bash
# final check at HEAD: the whole GPU suite, smoke(), and the bench with phase / kernel times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_gpu_tests_head.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3 | tee -a gpurun_out/r02_gpu_tests_head.log
timeout 900 python bench.py --steps 5 --warmup 3 --phase-times --kernel-times > gpurun_out/r02_bench_head.log 2> gpurun_out/r02_bench_head.err
tail -c 2500 gpurun_out/r02_bench_head.log; tail -4 gpurun_out/r02_bench_head.err | cut -c1-1200
