#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
for mode in 1 0; do
  echo "## N=2, 2-CTA GEMM, T1_ZERO_OVERLAP=$mode" | tee -a gpurun_out/r02_n2_2cta.log
  T1_ZERO_OVERLAP=$mode timeout 240 $TR bench.py --gpus 2 --steps 4 --warmup 3 --phase-times 2>> gpurun_out/r02_n2_2cta.log | tail -1 | cut -c1-200 | tee -a gpurun_out/r02_n2_2cta.log
done
echo "## N=2 cfg4 balanced" | tee -a gpurun_out/r02_n2_2cta.log
timeout 240 $TR bench.py --gpus 2 --steps 3 --warmup 3 --grid timerft --phase-times 2>> gpurun_out/r02_n2_2cta.log | tail -1 | cut -c1-200 | tee -a gpurun_out/r02_n2_2cta.log
grep "phase ms" gpurun_out/r02_n2_2cta.log
