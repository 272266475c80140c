#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1
echo "suite rc=$?"; tail -4 gpurun_out/gpu_tests.log
timeout 330 python bench.py --steps 5 --warmup 3 --phase-times --kernel-times > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-250; grep "phase ms" gpurun_out/bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
