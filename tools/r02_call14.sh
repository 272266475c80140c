#!/bin/bash
mkdir -p gpurun_out
# (1) decode-step ncu capture, small: gather, norm, then 2.4 layers of the chain (stream GEMM plain / fold / scaled + attention)
timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/r02_decode_full -c 14 python tools/decode_ncu.py 2 stream 1 > /dev/null 2>&1
ls -la gpurun_out/r02_decode_full.ncu-rep
# (2) decode attention cluster size vs context length
bash tools/r02_call13.sh
# (3) default bench exactly as the driver runs it (gpu_reference leg included), then the reference arm
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_default.log 2> gpurun_out/r02_bench_default.err ) 2>&1 | grep real
tail -1 gpurun_out/r02_bench_default.log | cut -c1-400
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.log 2>&1 ) 2>&1 | grep real
tail -1 gpurun_out/r02_bench_reference_arm.log | cut -c1-300
