#!/bin/bash
# N=1 dry run of BASELINE cfg3 / cfg4 / cfg5 (memory + save-level fallbacks) before the 8-GPU call
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference --phase-times"
echo "## cfg3 slots 32" | tee gpurun_out/r02_cfg345_n1.log
timeout 500 $B --slots 32 2>> gpurun_out/r02_cfg345_n1.log | tail -1 | tee -a gpurun_out/r02_cfg345_n1.log | cut -c1-300
echo "## cfg4 timerft + filter pass" | tee -a gpurun_out/r02_cfg345_n1.log
timeout 500 $B --grid timerft --filter-pass 8 2>> gpurun_out/r02_cfg345_n1.log | tail -1 | tee -a gpurun_out/r02_cfg345_n1.log | cut -c1-300
echo "## cfg5 slots 64" | tee -a gpurun_out/r02_cfg345_n1.log
timeout 700 $B --slots 64 2>> gpurun_out/r02_cfg345_n1.log | tail -1 | tee -a gpurun_out/r02_cfg345_n1.log | cut -c1-300
grep -E "phase ms|Error|error" gpurun_out/r02_cfg345_n1.log | cut -c1-300
