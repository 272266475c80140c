#!/bin/bash
# r02 call 1: golden GPU test, HF GPU reference arm, softmax poly-exp2 A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_compute_loss_golden_gpu.py -x -q 2>&1 | tail -5 | tee gpurun_out/r02_golden_gpu.log
nvidia-smi --query-gpu=name,memory.total --format=csv | tee gpurun_out/r02_smi.log
timeout 900 python bench.py --impl hf_gpu --steps 2 --warmup 1 > gpurun_out/r02_hf_gpu.log 2> gpurun_out/r02_hf_gpu.err
echo "hf_gpu rc=$?"; tail -1 gpurun_out/r02_hf_gpu.log; tail -5 gpurun_out/r02_hf_gpu.err
bash tools/ab_softmax_poly_exp2.sh
