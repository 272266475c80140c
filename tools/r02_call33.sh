#!/bin/bash
# N-invariance of the fix_vit=False path on NCCL: the same global batch (8 prompt-groups, two optimizer steps) at N = 1 and N = 2
mkdir -p gpurun_out
timeout 400 python bench.py --verify gpurun_out/verify_trainvit_n1.pt --train-vit 2>&1 | grep -a verify | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --verify gpurun_out/verify_trainvit_n2.pt --train-vit 2>&1 | grep -a verify | cut -c1-300
python tools/verify_compare.py gpurun_out/verify_trainvit_n1.pt gpurun_out/verify_trainvit_n2.pt | tee gpurun_out/r02_verify_train_vit_n1_vs_n2.txt
rm -f gpurun_out/verify_trainvit_n*.pt; exit 0
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep -a '^{' | cut -c1-400 | tee gpurun_out/r02_bench_2gpu_head.log
