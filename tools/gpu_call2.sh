#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_answer_scan_gpu.py -x -q > gpurun_out/scan_tests.log 2>&1; SCAN_RC=$?
tail -3 gpurun_out/scan_tests.log
for S in 8 16 4; do
  T1_DECODE_CLUSTER=$S T1_BENCH_MAX_CTAS=148 timeout 200 python tools/decode_bench.py 2 200 2>&1 | tail -1 | sed "s/^/cluster=$S /" | tee -a gpurun_out/decode_cluster.log
done
ROUTE=device; [ $SCAN_RC -ne 0 ] && ROUTE=host
timeout 330 python bench.py --reward-route $ROUTE --phase-times --kernel-times > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-300; grep "phase ms" gpurun_out/bench.err
