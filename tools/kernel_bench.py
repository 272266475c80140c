"""Per-kernel micro-benchmarks (CUDA events, L2 flushed between iterations). Usage: python tools/kernel_bench.py gemm"""
import json
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from time_r1_b200 import ops  # noqa: E402


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def bench_gemm():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    shapes = [(5836, 4608, 3584), (5836, 3584, 3584), (5836, 37888, 3584), (5836, 3584, 18944), (1600, 152064, 3584),
              (16384, 1280, 1176), (16384, 3840, 1280), (16384, 6912, 1280), (8192, 8192, 8192), (37888, 3584, 5840)]
    for M, N, K in shapes:
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: ops.gemm(a, b, out=out), flush=flush)
        ms_t = timeit(lambda: torch.matmul(a, b.t(), out=out), flush=flush)
        fl = 2.0 * M * N * K
        print(json.dumps({"kernel": "gemm_bf16_tn", "M": M, "N": N, "K": K, "ms": round(ms, 4),
                          "tflops": round(fl / ms / 1e9, 1), "cublas_ms": round(ms_t, 4),
                          "cublas_tflops": round(fl / ms_t / 1e9, 1)}), flush=True)
        del a, b, out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
    {"gemm": bench_gemm}[which]()
