#!/bin/bash
# One gpurun call: new-kernel tests -> reward-path microbench -> bench -> the rest of the GPU suite, inside a deadline.
DEADLINE=${1:-780}
START=$(date +%s)
left() { echo $(( DEADLINE - ($(date +%s) - START) )); }
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_answer_scan_gpu.py -x -q > gpurun_out/scan_tests.log 2>&1
SCAN_RC=$?
tail -3 gpurun_out/scan_tests.log
timeout 90 python tools/scan_bench.py > gpurun_out/scan_bench.log 2>&1; tail -2 gpurun_out/scan_bench.log
ROUTE=device; [ $SCAN_RC -ne 0 ] && ROUTE=host
timeout 330 python bench.py --reward-route $ROUTE --phase-times --kernel-times > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-600; grep "phase ms" gpurun_out/bench.err
L=$(left); echo "left for the full suite: $L s"
if [ $L -gt 60 ]; then
  timeout $L python -m pytest tests -m gpu -x -q --deselect tests/test_answer_scan_gpu.py --durations=15 > gpurun_out/gpu_tests.log 2>&1
  echo "full suite rc=$?"; tail -25 gpurun_out/gpu_tests.log
fi
