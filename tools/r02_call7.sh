#!/bin/bash
# 2 GPUs: two-step N-invariance with the lazy all-gather, bench A/B of T1_LAZY_GATHER
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 300 python bench.py --verify gpurun_out/verify_n1.pt 2>&1 | tail -1
timeout 300 $TR bench.py --gpus 2 --verify gpurun_out/verify_n2.pt 2>&1 | tail -1
python tools/verify_compare.py gpurun_out/verify_n1.pt gpurun_out/verify_n2.pt | tee gpurun_out/r02_verify2_n1_vs_n2.txt
rm -f gpurun_out/verify_n1.pt gpurun_out/verify_n2.pt
for mode in 1 0; do
  echo "## T1_LAZY_GATHER=$mode" | tee -a gpurun_out/r02_lazy_gather_n2.log
  T1_LAZY_GATHER=$mode timeout 400 $TR bench.py --gpus 2 --steps 4 --warmup 3 --phase-times 2>> gpurun_out/r02_lazy_gather_n2.log | tail -1 | cut -c1-300 | tee -a gpurun_out/r02_lazy_gather_n2.log
done
grep "phase ms" gpurun_out/r02_lazy_gather_n2.log
