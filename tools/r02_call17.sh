#!/bin/bash
# 8 x B200 at the shipped kernels (2-CTA GEMM, tcgen05 ViT attention, overlapped reduce-scatter, balanced cfg4 deal)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
B="bench.py --gpus 8 --no-cpu-baseline --no-gpu-reference --phase-times"
L=gpurun_out/r02_8gpu_final.log
: > $L
run() { echo "## $1" | tee -a $L; shift; timeout 240 "$@" 2>> $L | tail -1 | tee -a $L | cut -c1-200; }
run "cfg2 N=8" $TR $B --steps 4 --warmup 3
run "cfg3 N=8 (32 slots)" $TR $B --steps 3 --warmup 3 --slots 32
run "cfg4 N=8 (TimeRFT grids, balanced deal, filter pass)" $TR $B --steps 3 --warmup 3 --grid timerft --filter-pass 8
run "cfg5 N=8 (64 slots)" $TR $B --steps 2 --warmup 3 --slots 64
grep -E "phase ms|##" $L | cut -c1-300
