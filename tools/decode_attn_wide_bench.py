"""Decode attention, cluster form (8 CTAs per (group, KV head), DSMEM merge) vs wide form (every SM, L2 workspace merge):
one launch alone with L2 flushed / warm, at the prompt lengths of BASELINE cfg2 / cfg3 / cfg5 (2 prompt-groups, G = 8,
step 100).  Also checks that both forms give the same output on the timed inputs.
Usage: python tools/decode_attn_wide_bench.py [trace]   (trace: build with T1_EXTRA_NVCC=-DT1_DECODE_DBG)"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from time_r1_b200 import lib, ops  # noqa: E402
from time_r1_b200.lib import call, ptr, stream_ptr  # noqa: E402

DEV = "cuda"
BF16 = torch.bfloat16
G, Hq, Hkv, hd, cpad = 8, 28, 4, 128, 256


def rnd(*shape, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(*shape, device=DEV, generator=g).to(BF16)


def setup(Ps):
    n = len(Ps)
    ld = (Hq + 2 * Hkv) * hd
    raw = rnd(n * G, ld, seed=1)
    groups, rec = [], np.zeros((n, 7), np.int64)
    for gi, P in enumerate(Ps):
        S = P + G * cpad
        d = dict(K=rnd(S, Hkv * hd, seed=10 + gi), V=rnd(S, Hkv * hd, seed=20 + gi), cos=torch.rand(cpad, hd // 2, device=DEV),
                 sin=torch.rand(cpad, hd // 2, device=DEV))
        groups.append(d)
        rec[gi] = (d["K"].data_ptr(), d["V"].data_ptr(), 0, 0, d["cos"].data_ptr(), d["sin"].data_ptr(), P)
    maps = ops.kv_cache_tensor_maps([t for g in groups for t in (g["K"], g["V"])], Hkv, hd)
    return raw, groups, torch.from_numpy(rec).to(DEV), ld, maps


flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, flush_l2=True, reps=25):
    ts = []
    for i in range(reps):
        if flush_l2:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def graph_time(fn, n=28, reps=10):
    """n back-to-back launches captured in one graph (PDL-chained like the decode step's), per launch"""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    step = 100
    for name, Ps in (("cfg2", [4216, 4216]), ("cfg3", [8312, 8312]), ("cfg5", [16504, 16504]), ("cfg2_1group", [4216])):
        raw, groups, tab, ld, maps = setup(Ps)
        n = len(Ps)
        st = torch.tensor([step], device=DEV, dtype=torch.int32)
        kv_bytes = sum(2.0 * Hkv * hd * 2 * (P + G * step) for P in Ps)
        outs = {}
        forms = [("cluster", 8)] + [("wide", s) for s in (0, 8, 12)]
        for form, arg in forms:
            out = torch.empty(n * G, Hq * hd, device=DEV, dtype=BF16)
            if form == "cluster":
                def fn(o=out):
                    call("t1_decode_attn_cluster", ptr(raw), ld, ptr(tab), maps.ctypes.data, n, G, Hq, Hkv, hd, cpad, ptr(st), ptr(o),
                         o.stride(0), hd ** -0.5, arg, stream_ptr())
                splits = arg
            else:
                if arg * Hkv * n > 148:
                    continue
                splits, ws = ops.decode_attn_wide_workspace(n, Hkv, DEV, arg)

                def fn(o=out, splits=splits, ws=ws):
                    call("t1_decode_attn_wide", ptr(raw), ld, ptr(tab), maps.ctypes.data, n, G, Hq, Hkv, hd, cpad, ptr(st), ptr(o),
                         o.stride(0), hd ** -0.5, splits, ptr(ws), ws.numel(), stream_ptr())
            fn()
            torch.cuda.synchronize()
            outs[(form, arg)] = out.clone()
            us = timeit(fn) * 1e3
            us_warm = timeit(fn, flush_l2=False) * 1e3
            lib.load().t1_set_pdl(1)
            us_graph = graph_time(fn) * 1e3
            print(json.dumps({"shape": name, "P": Ps, "step": step, "form": form, "ctas_per_group_head": splits, "ctas": splits * Hkv * n,
                              "us_l2_flushed": round(us, 2), "gbs": round(kv_bytes / us / 1e3, 1),
                              "frac_of_6549.8": round(kv_bytes / us / 1e3 / 6549.8, 3), "us_warm_l2": round(us_warm, 2),
                              "us_in_graph_x28_same_cache": round(us_graph, 2), "unique_kv_mb": round(kv_bytes / 1e6, 2)}), flush=True)
        ref = outs[("cluster", 8)].float()
        for k, o in outs.items():
            err = (o.float() - ref).abs().max().item()
            assert err < 2e-2, (k, err)
        print(json.dumps({"shape": name, "max_abs_diff_vs_cluster": {f"{k[0]}{k[1]}": round((o.float() - ref).abs().max().item(), 5) for k, o in outs.items()}}), flush=True)

    if "trace" in sys.argv:
        L = lib.load()
        rd = L.t1_decode_attn_dbg_read
        rd.argtypes = [ctypes.c_void_p]
        rd.restype = ctypes.c_int
        Ps = [4216, 4216]
        raw, groups, tab, ld, maps = setup(Ps)
        st = torch.tensor([step], device=DEV, dtype=torch.int32)
        out = torch.empty(len(Ps) * G, Hq * hd, device=DEV, dtype=BF16)
        splits, ws = ops.decode_attn_wide_workspace(len(Ps), Hkv, DEV, 0)
        for _ in range(3):
            call("t1_decode_attn_wide", ptr(raw), ld, ptr(tab), maps.ctypes.data, len(Ps), G, Hq, Hkv, hd, cpad, ptr(st), ptr(out),
                 out.stride(0), hd ** -0.5, splits, ptr(ws), ws.numel(), stream_ptr())
        torch.cuda.synchronize()
        buf = np.zeros(1024, np.int64)
        assert rd(buf.ctypes.data) == 0
        t0 = buf[900]
        print("wide S", splits, "CTA 0 marks (clk rel start): after_qbuild", buf[901] - t0, "loop_end", buf[902] - t0, "pre_store", buf[903] - t0,
              "stored", buf[904] - t0, "peers_arrived", buf[905] - t0, "post_sync", buf[906] - t0, "merged", buf[907] - t0)
        for it in range(4):
            r = buf[it * 8: it * 8 + 6] - t0
            print(f"  it {it:2d} tma_issued {r[0]:7d} mma_full {r[1]:7d} mma_pready {r[2]:7d} sm_sfull {r[3]:7d} sm_xch {r[4]:7d} sm_pdone {r[5]:7d}")


if __name__ == "__main__":
    main()
