#!/bin/bash
# (1) decode step: wide vs cluster decode attention at the cfg3 / cfg5 context lengths; (2) reference-model prompt prefill beside the decode (side_ctas) at the shipped kernels
mkdir -p gpurun_out
out=gpurun_out/r02_decode_wide_vs_cluster_by_context.txt
: > $out
for nv in 8192 16384; do
  echo "## vision tokens $nv" | tee -a $out
  T1_BENCH_NV=$nv T1_BENCH_VARIANTS='[{"decode_attn": "cluster"}, {"decode_attn": "wide"}, {"decode_attn": "cluster"}, {"decode_attn": "wide"}]' \
    timeout 300 python tools/decode_bench.py 2 200 2>&1 | tail -4 | cut -c1-260 | tee -a $out
done
out=gpurun_out/r02_side_ctas_sweep.txt
: > $out
for sc in 0 16 24 32; do
  echo "## T1_SIDE_CTAS=$sc T1_DECODE_CTAS=$((148 - sc))" | tee -a $out
  T1_SIDE_CTAS=$sc T1_DECODE_CTAS=$((148 - sc)) timeout 300 python bench.py --steps 3 --warmup 3 --no-gpu-reference --no-cpu-baseline --phase-times 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(json.dumps({k: d[k] for k in ('value', 'ms_per_step')}))
    elif l.startswith('[phase'):
        print(l.strip())
" | tee -a $out
done
