#!/bin/bash
# decode attention cluster size vs context length (cfg2 / cfg3 / cfg5 prompt lengths)
mkdir -p gpurun_out
OUT=gpurun_out/r02_decode_cluster_vs_context.txt
: > $OUT
for nv in 4096 8192 16384; do
  for S in 8 16; do
    echo "## vision tokens $nv, T1_DECODE_CLUSTER=$S" | tee -a $OUT
    T1_BENCH_NV=$nv T1_DECODE_CLUSTER=$S T1_BENCH_MAX_CTAS=148 timeout 200 python tools/decode_bench.py 2 200 2>&1 | tail -1 | tee -a $OUT
  done
done
