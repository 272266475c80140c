#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_multigroup_gpu.py tests/test_gemm_gpu.py -m gpu -x -q 2>&1 | tail -5
out=gpurun_out/r02_side_ctas_sweep_v2.txt
: > $out
for sc in 16 24 0; do
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
