"""RMSNorm forward at the step's shapes: us per launch and HBM GB/s (bytes = x + delta read, x_out + y written).
L2 flushed between launches, CUDA events, median of 9.   python tools/norm_bench.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from time_r1_b200 import ops  # noqa: E402

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for T, D, with_delta in [(5888, 3584, True), (4216, 3584, True), (1672, 3584, True), (16384, 1280, True), (16384, 1280, False), (5888, 3584, False)]:
    x = torch.randn(T, D, device=dev).to(torch.bfloat16)
    d = torch.randn(T, D, device=dev).to(torch.bfloat16) if with_delta else None
    w = torch.ones(D, device=dev, dtype=torch.bfloat16)
    out, xo = torch.empty_like(x), torch.empty_like(x)
    ts = []
    for i in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.rmsnorm(x, w, 1e-6, delta=d, out=out, x_out=xo if with_delta else None)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    by = T * D * 2 * (4 if with_delta else 2)
    print(json.dumps({"kernel": "rmsnorm_fwd", "T": T, "D": D, "delta": with_delta, "us": round(ms * 1e3, 1), "gbs": round(by / ms / 1e6, 1)}))
