#!/bin/bash
# wide decode attention: parity, stand-alone timing vs the cluster form, decode-step A/B, CTA trace
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "decode_attention" 2>&1 | tail -15 | tee gpurun_out/r02_wide_tests.log
timeout 300 python tools/decode_attn_wide_bench.py 2>&1 | tee gpurun_out/r02_decode_attn_wide_vs_cluster.txt | cut -c1-400
for form in cluster wide; do
  echo "## T1_DECODE_ATTN=$form" | tee -a gpurun_out/r02_decode_step_wide_ab.txt
  T1_DECODE_ATTN=$form T1_BENCH_MAX_CTAS=0 timeout 300 python tools/decode_bench.py 2 200 2>&1 | tail -2 | tee -a gpurun_out/r02_decode_step_wide_ab.txt
done
touch time_r1_b200/csrc/decode_attn_tc.cu
T1_EXTRA_NVCC=-DT1_DECODE_DBG timeout 300 python -m time_r1_b200.build 2>&1 | tail -1
timeout 200 python tools/decode_attn_wide_bench.py trace 2>&1 | grep -A6 "^wide S" | tee gpurun_out/r02_decode_attn_wide_trace.txt
