"""ViT attention (head_dim 80, Qwen2.5-VL-7B: 16 heads, cfg2 = 16 temporal slots of 32x32 patches = 16384 tokens):
tcgen05 kernel (t1_attn_fwd_tc<80>) vs the HMMA kernel it replaces vs flash_attn_varlen_func (FA2 recompiled for sm_100,
what the reference runs — SURVEY §2.3's bar), windowed (64-token windows) and full (1024-token slots) blocks.
CUDA events, L2 flushed between launches, median of 9; parity of the three against an fp32 torch reference on one slot.
    python tools/vit_attn_bench.py  ->  one JSON line per (kernel, mode)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from time_r1_b200 import config, indexing, ops  # noqa: E402

dev = "cuda"
cfg = config.PRESETS["qwen2.5-vl-7b"]()
H, hd, vd = cfg.vit_heads, cfg.vit_head_dim, cfg.vit_hidden
grid = [(16, 32, 32)]
plan = indexing.vision_plan(cfg, grid)
N = plan.n_patches
g = torch.Generator(device=dev).manual_seed(0)
qkv = (torch.randn(N, 3 * vd, device=dev, generator=g) * 0.5).to(torch.bfloat16)
i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
ranges = {"window64": (i32(plan.win_lo), i32(plan.win_hi)), "full1024": (i32(plan.full_lo), i32(plan.full_hi))}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
scale = hd ** -0.5


def timed(fn):
    ts = []
    for i in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def cu_of(lo, hi):
    lo, hi = lo.cpu().numpy(), hi.cpu().numpy()
    starts = np.unique(lo)
    return torch.from_numpy(np.concatenate([starts, [hi.max()]]).astype(np.int32)).to(dev)


try:
    from flash_attn import flash_attn_varlen_func
except Exception as e:      # pragma: no cover
    flash_attn_varlen_func = None
    print(json.dumps({"flash_attn": f"unavailable: {e}"}))

for mode, (lo, hi) in ranges.items():
    q, k, v = qkv[:, :vd], qkv[:, vd:2 * vd], qkv[:, 2 * vd:]
    keys_per_row = float((hi - lo).float().mean().item())
    flops = 4.0 * N * H * hd * keys_per_row
    out_tc = torch.empty((N, vd), device=dev, dtype=torch.bfloat16)
    out_hm = torch.empty_like(out_tc)
    ops.attn_fwd_tc(q, k, v, lo, hi, 0, H, H, hd, scale, out=out_tc, need_lse=False)
    ops.attn_fwd(q, k, v, lo, hi, 0, H, H, hd, scale, out=out_hm, need_lse=False)
    # fp32 reference on the first 2048 rows
    R = 2048
    qf, kf, vf = (t[:R].float().view(R, H, hd).transpose(0, 1) for t in (q, k, v))
    sc = torch.einsum("hid,hjd->hij", qf, kf) * scale
    jj = torch.arange(R, device=dev)
    vis = (jj[None, :] >= lo[:R, None]) & (jj[None, :] < hi[:R, None])
    ref = torch.einsum("hij,hjd->ihd", torch.softmax(sc.masked_fill(~vis[None], float("-inf")), -1), vf).reshape(R, vd)
    err = {"tc": float((out_tc[:R].float() - ref).abs().max()), "hmma": float((out_hm[:R].float() - ref).abs().max())}
    res = {}
    if mode == "window64":
        out_w = torch.zeros_like(out_tc)
        ops.attn_fwd_tc_win(q, k, v, lo, hi, H, hd, scale, out=out_w)
        err["tc_win"] = float((out_w[:R].float() - ref).abs().max())
        err["tc_win_vs_hmma_all_rows"] = float((out_w.float() - out_hm.float()).abs().max())
        res["tcgen05_win (persistent, block-local)"] = timed(lambda: ops.attn_fwd_tc_win(q, k, v, lo, hi, H, hd, scale, out=out_w))
    res.update({"tcgen05": timed(lambda: ops.attn_fwd_tc(q, k, v, lo, hi, 0, H, H, hd, scale, out=out_tc, need_lse=False)),
           "hmma": timed(lambda: ops.attn_fwd(q, k, v, lo, hi, 0, H, H, hd, scale, out=out_hm, need_lse=False))})
    if flash_attn_varlen_func is not None:
        cu = cu_of(lo, hi)
        mx = int((cu[1:] - cu[:-1]).max())
        q3, k3, v3 = (t.reshape(N, H, hd).contiguous() for t in (q, k, v))
        o_fa = flash_attn_varlen_func(q3, k3, v3, cu, cu, mx, mx, softmax_scale=scale, causal=False)
        err["fa2"] = float((o_fa[:R].reshape(R, vd).float() - ref).abs().max())
        res["fa2_varlen(contiguous q/k/v)"] = timed(lambda: flash_attn_varlen_func(q3, k3, v3, cu, cu, mx, mx, softmax_scale=scale, causal=False))
    for name, ms in res.items():
        print(json.dumps({"mode": mode, "kernel": name, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1), "tokens": N, "heads": H,
                          "head_dim": hd, "keys_per_row": keys_per_row}))
    print(json.dumps({"mode": mode, "max_abs_err_vs_fp32": err}))
