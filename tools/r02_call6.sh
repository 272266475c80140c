#!/bin/bash
# 8 x B200: N-invariance at N=8, BASELINE cfg2 (overlapped vs blocking ZeRO-2 exchange), cfg3, cfg4 (+ filter pass), cfg5
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
B="bench.py --gpus 8 --no-cpu-baseline --no-gpu-reference --phase-times"
L=gpurun_out/r02_8gpu.log
: > $L
timeout 200 python bench.py --verify gpurun_out/verify_n1.pt 2>&1 | tail -1 | tee -a $L
timeout 300 $TR bench.py --gpus 8 --verify gpurun_out/verify_n8.pt 2>&1 | tail -1 | tee -a $L
python tools/verify_compare.py gpurun_out/verify_n1.pt gpurun_out/verify_n8.pt | tee gpurun_out/r02_verify_n1_vs_n8.txt
rm -f gpurun_out/verify_n1.pt gpurun_out/verify_n8.pt
run() { echo "## $1" | tee -a $L; shift; timeout 420 env "$@" 2>> $L | tail -1 | tee -a $L | cut -c1-260; }
run "cfg2 N=8 overlapped exchange (default)" T1_X=1 $TR $B --steps 4 --warmup 3
run "cfg2 N=8 blocking exchange" T1_ZERO_OVERLAP=0 $TR $B --steps 4 --warmup 3
run "cfg3 N=8 (32 slots)" T1_X=1 $TR $B --steps 3 --warmup 3 --slots 32
run "cfg4 N=8 (TimeRFT grids + filter pass)" T1_X=1 $TR $B --steps 3 --warmup 3 --grid timerft --filter-pass 8
run "cfg5 N=8 (64 slots)" T1_X=1 $TR $B --steps 2 --warmup 3 --slots 64
grep -E "phase ms|##" $L | cut -c1-300
