"""Bring-up tool: cluster decode attention — mismatching rows vs the stand-alone rope/append, then timing per cluster size."""
import os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from time_r1_b200 import ops
from time_r1_b200.lib import call, ptr, stream_ptr
DEV = "cuda"
BF16 = torch.bfloat16

def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(*shape, device=DEV, generator=g) * scale).to(BF16)

def setup(G, Hq, Hkv, Ps, cpad):
    hd = 128
    n = len(Ps)
    ld = (Hq + 2 * Hkv) * hd
    raw = rnd(n * G, ld, seed=1)
    groups, rec = [], np.zeros((n, 7), np.int64)
    for gi, P in enumerate(Ps):
        S = P + G * cpad
        d = dict(K=rnd(S, Hkv * hd, seed=10 + gi), V=rnd(S, Hkv * hd, seed=20 + gi), cos=torch.rand(cpad, hd // 2, device=DEV), sin=torch.rand(cpad, hd // 2, device=DEV))
        nch = (P + 127) // 128 + G * (cpad // 128)
        d["ws_o"] = torch.empty(nch * Hkv * 64 * 128, device=DEV); d["ws_ml"] = torch.empty(nch * Hkv * 64 * 2, device=DEV)
        groups.append(d)
        rec[gi] = (d["K"].data_ptr(), d["V"].data_ptr(), d["ws_o"].data_ptr(), d["ws_ml"].data_ptr(), d["cos"].data_ptr(), d["sin"].data_ptr(), P)
    maps = ops.kv_cache_tensor_maps([t for g in groups for t in (g["K"], g["V"])], Hkv, hd)
    return raw, groups, torch.from_numpy(rec).to(DEV), ld, maps

if "check" in sys.argv:
    G, Hq, Hkv, Ps, step, cpad = 8, 28, 4, [1000, 4216], 130, 256
    hd = 128
    for cluster in (1, 8):
        raw, groups, tab, ld, maps = setup(G, Hq, Hkv, Ps, cpad)
        st = torch.tensor([step], device=DEV, dtype=torch.int32)
        ref_qkv = raw.clone()
        refK, refV = [g["K"].clone() for g in groups], [g["V"].clone() for g in groups]
        for gi, P in enumerate(Ps):
            call("t1_decode_rope_append", ptr(ref_qkv[gi * G:(gi + 1) * G]), ld, G, Hq, Hkv, hd, ptr(groups[gi]["cos"]), ptr(groups[gi]["sin"]), ptr(st), ptr(refK[gi]), ptr(refV[gi]), P, cpad, stream_ptr())
        out = torch.empty(len(Ps) * G, Hq * hd, device=DEV, dtype=BF16)
        call("t1_decode_attn_cluster", ptr(raw), ld, ptr(tab), maps.ctypes.data, len(Ps), G, Hq, Hkv, hd, cpad, ptr(st), ptr(out), out.stride(0), hd ** -0.5, cluster, stream_ptr())
        torch.cuda.synchronize()
        for gi, P in enumerate(Ps):
            for nm, a, b in (("K", groups[gi]["K"], refK[gi]), ("V", groups[gi]["V"], refV[gi])):
                bad = (a != b).any(1).nonzero().flatten().tolist()
                if bad:
                    r = bad[0]
                    cols = (a[r] != b[r]).nonzero().flatten().tolist()
                    print(json.dumps({"cluster": cluster, "group": gi, "which": nm, "bad_rows": bad[:12], "n_bad": len(bad), "expected_rows": [P + b_ * cpad + step for b_ in range(G)][:4],
                                      "row": r, "n_cols": len(cols), "cols": cols[:8], "got": a[r, cols[:4]].float().tolist(), "want": b[r, cols[:4]].float().tolist()}))
        print("checked cluster", cluster, flush=True)

# timing at cfg2: 2 groups, P=4216, G=8, step 100
G, Hq, Hkv, Ps, step, cpad = 8, 28, 4, [4216, 4216], 100, 256
hd = 128
raw, groups, tab, ld, maps = setup(G, Hq, Hkv, Ps, cpad)
st = torch.tensor([step], device=DEV, dtype=torch.int32)
out = torch.empty(len(Ps) * G, Hq * hd, device=DEV, dtype=BF16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
kv_bytes = sum(2.0 * Hkv * hd * 2 * (P + G * (step + 1)) for P in Ps) + 2.0 * len(Ps) * G * Hq * hd * 2
def timeit(fn, flush_l2=True):
    ts = []
    for i in range(15):
        if flush_l2: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]
for cluster in (4, 8, 16):
    fn = lambda: call("t1_decode_attn_cluster", ptr(raw), ld, ptr(tab), maps.ctypes.data, len(Ps), G, Hq, Hkv, hd, cpad, ptr(st), ptr(out), out.stride(0), hd ** -0.5, cluster, stream_ptr())
    ms = timeit(fn)
    ms_warm = timeit(fn, flush_l2=False)
    print(json.dumps({"kernel": "cluster", "S": cluster, "us": round(ms * 1e3, 2), "gbs": round(kv_bytes / ms / 1e6, 1), "us_warm_l2": round(ms_warm * 1e3, 2)}), flush=True)
fn = lambda: call("t1_decode_attn_groups", ptr(raw), ld, ptr(tab), len(Ps), G, Hq, Hkv, hd, max(Ps), cpad, ptr(st), ptr(out), out.stride(0), hd ** -0.5, 1, stream_ptr())
ms = timeit(fn)
print(json.dumps({"kernel": "split+combine", "us": round(ms * 1e3, 2), "gbs": round(kv_bytes / ms / 1e6, 1)}), flush=True)
