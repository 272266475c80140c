"""Finishing-time spread of the stream GEMM's CTAs (build with T1_EXTRA_NVCC=-DT1_SG_DBG)."""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from time_r1_b200 import lib, ops
L = lib.load()
fn = L.t1_stream_gemm_cta_times; fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (M, N, K) in [(16, 37888, 3584), (16, 3584, 18944), (16, 3584, 3584), (16, 152064, 3584)]:
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    ws = ops.stream_gemm_workspace(M, "cuda"); out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    res = []
    for it in range(5):
        flush.zero_(); torch.cuda.synchronize()
        ops.stream_gemm(x, w, ws, out=out); torch.cuda.synchronize()
        buf = np.zeros(3 * 512, np.uint64); assert fn(buf.ctypes.data) == 0
        t = buf.reshape(3, 512).astype(np.int64)
        n = min(148, int((t[2] > 0).sum()))
        st, en = t[0, :n], t[2, :n]
        t0 = st.min()
        res.append((float((en - t0).min()) / 1e3, float(np.median(en - t0)) / 1e3, float((en - t0).max()) / 1e3, float((st - t0).max()) / 1e3))
    r = res[-1]
    order = np.argsort(en)
    print(json.dumps({"M": M, "N": N, "K": K, "first_done_us": round(r[0], 2), "median_done_us": round(r[1], 2), "last_done_us": round(r[2], 2),
                      "last_start_us": round(r[3], 2), "slowest_ctas": order[-8:].tolist(), "fastest_ctas": order[:8].tolist(),
                      "MB": round(2 * N * K / 1e6, 1), "gbs_at_last": round(2 * N * K / r[2] / 1e3, 1)}), flush=True)
