#!/bin/bash
# decode step at the cfg3 / cfg5 context lengths at the final kernels (decode_attn = "auto" -> wide form from 8192 prompt tokens on)
mkdir -p gpurun_out
out=gpurun_out/r02_decode_step_final_by_context.txt
: > $out
for nv in 8192 16384; do
  echo "## vision tokens $nv" | tee -a $out
  T1_BENCH_NV=$nv T1_BENCH_VARIANTS='[{}, {"decode_attn": "cluster"}]' timeout 300 python tools/decode_bench.py 2 200 2>&1 | tail -2 | cut -c1-230 | tee -a $out
done
