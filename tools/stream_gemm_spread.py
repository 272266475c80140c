"""Finishing-time spread of the stream GEMM's CTAs (build with T1_EXTRA_NVCC=-DT1_SG_DBG)."""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from time_r1_b200 import lib, ops
L = lib.load()
fn = L.t1_stream_gemm_cta_times; fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
hist = []
rel_blk = {}
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
        rel_blk.setdefault((M, N, K), []).append((en - t0) / np.median(en - t0))
        res.append((float((en - t0).min()) / 1e3, float(np.median(en - t0)) / 1e3, float((en - t0).max()) / 1e3, float((st - t0).max()) / 1e3))
    r = res[-1]
    smid = t[1, :n]
    dur = (en - t0).astype(np.float64)
    hist.append((smid.copy(), dur.copy()))
    order = np.argsort(en)
    print(json.dumps({"M": M, "N": N, "K": K, "first_done_us": round(r[0], 2), "median_done_us": round(r[1], 2), "last_done_us": round(r[2], 2),
                      "last_start_us": round(r[3], 2), "slowest_ctas": order[-8:].tolist(), "fastest_ctas": order[:8].tolist(),
                      "MB": round(2 * N * K / 1e6, 1), "gbs_at_last": round(2 * N * K / r[2] / 1e3, 1)}), flush=True)

# is the slowness a property of the SM?  correlate per-smid finishing times across shapes / launches
import itertools
by = []
for smid, dur in hist:
    v = np.full(256, np.nan); v[smid] = dur / np.median(dur); by.append(v)
by = np.array(by)
ok = ~np.isnan(by).any(0)
print("distinct smids per launch:", [int((~np.isnan(b)).sum()) for b in by], "blockIdx==smid:", [bool((h[0] == np.arange(len(h[0]))).all()) for h in hist])
for a, b in itertools.combinations(range(len(by)), 2):
    print("corr of relative finishing time by smid, launches", a, b, round(float(np.corrcoef(by[a][ok], by[b][ok])[0, 1]), 3))
m = np.nanmean(by, 0)
slow = np.argsort(-np.nan_to_num(m, nan=0))[:16]
print("slowest smids:", slow.tolist(), "rel:", np.round(m[slow], 3).tolist())
print("rel time by smid decile:", np.round(np.nanpercentile(m, [0, 10, 25, 50, 75, 90, 100]), 3).tolist())

# ... or of the block index (launch order)?  mean relative finishing time per block, binned by 16 blocks
for key, v in rel_blk.items():
    v = np.array([x for x in v if len(x) == 148])
    mb = v.mean(0)
    print(key, "rel time by block bin:", np.round([mb[i:i + 16].mean() for i in range(0, 148, 16)], 3).tolist(),
          "corr launch0/launch4 by block:", round(float(np.corrcoef(v[0], v[-1])[0, 1]), 3))
print("smid of blocks 0..15:", hist[0][0][:16].tolist(), " blocks 132..147:", hist[0][0][132:148].tolist())
