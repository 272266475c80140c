#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
timeout 300 python bench.py --verify gpurun_out/verify_n1.pt 2>&1 | tail -1
timeout 300 $TR bench.py --gpus 2 --verify gpurun_out/verify_n2.pt 2>&1 | tail -1
python tools/verify_compare.py gpurun_out/verify_n1.pt gpurun_out/verify_n2.pt | tee gpurun_out/r02_verify_final_n1_vs_n2.txt
rm -f gpurun_out/verify_n1.pt gpurun_out/verify_n2.pt
timeout 300 $TR bench.py --gpus 2 --steps 4 --warmup 3 2>/dev/null | tail -1 | cut -c1-200 | tee gpurun_out/r02_bench_2gpu_final.log
