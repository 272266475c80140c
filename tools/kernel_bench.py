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
              (1672, 3584, 3584), (1672, 3584, 18944), (1672, 4608, 3584), (4216, 3584, 18944), (16384, 1280, 1176), (16384, 3840, 1280), (16384, 6912, 1280), (8192, 8192, 8192), (37888, 3584, 5840)]
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


def bench_decode():
    from time_r1_b200.lib import call, ptr, stream_ptr
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for M, N, K in [(8, 4608, 3584), (8, 3584, 3584), (8, 37888, 3584), (8, 3584, 18944), (8, 152064, 3584), (32, 37888, 3584)]:
        x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: ops.skinny_gemm(x, w, out=out), flush=flush)
        ms_t = timeit(lambda: torch.matmul(x, w.t(), out=out), flush=flush)
        by = 2.0 * N * K
        print(json.dumps({"kernel": "skinny_gemm", "M": M, "N": N, "K": K, "ms": round(ms, 4), "gbs": round(by / ms / 1e6, 1),
                          "cublas_ms": round(ms_t, 4), "cublas_gbs": round(by / ms_t / 1e6, 1)}), flush=True)
    B, Hq, Hkv, hd, P, cpad = 8, 28, 4, 128, 4216, 256
    S = P + B * cpad
    Kc = torch.randn(S, Hkv * hd, device="cuda").to(torch.bfloat16)
    Vc = torch.randn(S, Hkv * hd, device="cuda").to(torch.bfloat16)
    q = torch.randn(B, Hq * hd, device="cuda").to(torch.bfloat16)
    n_chunks = (P + 127) // 128 + B * (cpad // 128)
    rows_pad = 64
    ws_o = torch.empty(n_chunks * Hkv * rows_pad * 128, device="cuda")
    ws_ml = torch.empty(n_chunks * Hkv * rows_pad * 2, device="cuda")
    out = torch.empty(B, Hq * hd, device="cuda", dtype=torch.bfloat16)
    for step in (0, 100, 199):
        st = torch.tensor([step], device="cuda", dtype=torch.int32)
        fn = lambda: call("t1_decode_attn", ptr(q), q.stride(0), ptr(Kc), ptr(Vc), B, Hq, Hkv, hd, P, cpad, ptr(st), ptr(ws_o), ptr(ws_ml), ptr(out), out.stride(0), hd ** -0.5, stream_ptr())
        ms = timeit(fn, flush=flush)
        ms_hot = timeit(fn, flush=None)
        by = 2.0 * Hkv * hd * 2 * (P + B * (step + 1))
        print(json.dumps({"kernel": "decode_attn+combine", "step": step, "ms_cold": round(ms, 4), "gbs_cold": round(by / ms / 1e6, 1), "ms_hot": round(ms_hot, 4)}), flush=True)
    for T, D in [(8, 3584), (5888, 3584)]:
        x = torch.randn(T, D, device="cuda").to(torch.bfloat16); w = torch.ones(D, device="cuda", dtype=torch.bfloat16); d = torch.randn(T, D, device="cuda").to(torch.bfloat16); y = torch.empty_like(x)
        ms = timeit(lambda: ops.rmsnorm(x, w, 1e-6, delta=d, out=y), flush=flush)
        print(json.dumps({"kernel": "rmsnorm+residual", "T": T, "D": D, "ms": round(ms, 4), "gbs": round(T * D * 2 * 4 / ms / 1e6, 1)}), flush=True)


def graph_time(fns, reps=5):
    """Capture the launches in `fns` into one CUDA graph and time its replay (in-pipeline cost, no launch gaps)."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    ts = []
    for _ in range(reps):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] / len(fns)


def bench_decode_graph():
    """Per-kernel cost inside a graph of 28 back-to-back launches cycling over distinct weight copies (> L2)."""
    from time_r1_b200.lib import call, ptr, stream_ptr
    NL = 28
    for M, N, K in [(8, 4608, 3584), (8, 3584, 3584), (8, 37888, 3584), (8, 3584, 18944), (8, 152064, 3584)]:
        ncopy = max(2, min(NL, int(300e6 // (2 * N * K)) + 1))
        ws = [torch.randn(N, K, device="cuda").to(torch.bfloat16) for _ in range(ncopy)]
        x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ms = graph_time([(lambda w=ws[i % ncopy]: ops.skinny_gemm(x, w, out=out)) for i in range(NL)])
        ms_t = graph_time([(lambda w=ws[i % ncopy]: torch.matmul(x, w.t(), out=out)) for i in range(NL)])
        by = 2.0 * N * K
        print(json.dumps({"kernel": "skinny_gemm(in-graph)", "M": M, "N": N, "K": K, "us": round(ms * 1e3, 2), "gbs": round(by / ms / 1e6, 1),
                          "cublas_us": round(ms_t * 1e3, 2), "cublas_gbs": round(by / ms_t / 1e6, 1)}), flush=True)
        del ws
    B, Hq, Hkv, hd, P, cpad = 8, 28, 4, 128, 4216, 256
    S = P + B * cpad
    kvs = [(torch.randn(S, Hkv * hd, device="cuda").to(torch.bfloat16), torch.randn(S, Hkv * hd, device="cuda").to(torch.bfloat16)) for _ in range(NL)]
    q = torch.randn(B, Hq * hd, device="cuda").to(torch.bfloat16)
    n_chunks = (P + 127) // 128 + B * (cpad // 128)
    ws_o = torch.empty(n_chunks * Hkv * 64 * 128, device="cuda"); ws_ml = torch.empty(n_chunks * Hkv * 64 * 2, device="cuda")
    out = torch.empty(B, Hq * hd, device="cuda", dtype=torch.bfloat16)
    st = torch.tensor([100], device="cuda", dtype=torch.int32)
    ms = graph_time([(lambda kv=kvs[i]: call("t1_decode_attn", ptr(q), q.stride(0), ptr(kv[0]), ptr(kv[1]), B, Hq, Hkv, hd, P, cpad, ptr(st), ptr(ws_o), ptr(ws_ml), ptr(out), out.stride(0), hd ** -0.5, stream_ptr())) for i in range(NL)])
    by = 2.0 * Hkv * hd * 2 * (P + B * 101)
    print(json.dumps({"kernel": "decode_attn+combine(in-graph)", "us": round(ms * 1e3, 2), "gbs": round(by / ms / 1e6, 1)}), flush=True)
    x = torch.randn(8, 3584, device="cuda").to(torch.bfloat16); w = torch.ones(3584, device="cuda", dtype=torch.bfloat16); d = torch.randn(8, 3584, device="cuda").to(torch.bfloat16); y = torch.empty_like(x)
    print(json.dumps({"kernel": "rmsnorm[8x3584](in-graph)", "us": round(graph_time([lambda: ops.rmsnorm(x, w, 1e-6, delta=d, out=y)] * NL) * 1e3, 2)}), flush=True)
    gu = torch.randn(8, 37888, device="cuda").to(torch.bfloat16); act = torch.empty(8, 18944, device="cuda", dtype=torch.bfloat16)
    print(json.dumps({"kernel": "swiglu[8x18944](in-graph)", "us": round(graph_time([lambda: ops.swiglu(gu, out=act)] * NL) * 1e3, 2)}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
    {"gemm": bench_gemm, "decode": bench_decode, "decode_graph": bench_decode_graph}[which]()
