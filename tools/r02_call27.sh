#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_trainer_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/decode_timeline.py 2 2>&1 | tee gpurun_out/r02_decode_timeline_fold.txt | tail -45
