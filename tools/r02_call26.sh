#!/bin/bash
# fix_vit=False: ViT-block gradients vs the oracle, head_dim-80 attention backward, then the cfg2 step with a trained tower
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_kernels_gpu.py -m gpu -x -q -k "vit or attention_backward or fix_vit or grpo_step" 2>&1 | tail -15 | tee gpurun_out/r02_trainvit_tests.log
timeout 600 python bench.py --train-vit --steps 3 --warmup 3 --no-gpu-reference --no-cpu-baseline --phase-times > gpurun_out/r02_bench_train_vit.log 2> gpurun_out/r02_bench_train_vit.err
tail -c 1500 gpurun_out/r02_bench_train_vit.log; tail -20 gpurun_out/r02_bench_train_vit.err
