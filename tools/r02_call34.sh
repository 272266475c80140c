#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_gemm_gpu.py -m gpu -x -q -k "fused or gemm or grpo_step" 2>&1 | tail -6
for f in 0 1; do
  echo "## T1_FUSED_HEAD=$f" | tee -a gpurun_out/r02_fused_head_ab.txt
  T1_FUSED_HEAD=$f timeout 300 python bench.py --steps 3 --warmup 3 --no-gpu-reference --no-cpu-baseline --phase-times 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(json.dumps({k: d[k] for k in ('value', 'ms_per_step', 'hbm_gb')}))
    elif l.startswith('[phase'):
        print(l.strip())
" | tee -a gpurun_out/r02_fused_head_ab.txt
done
