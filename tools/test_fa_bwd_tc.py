"""Bring-up / timing tool for the tcgen05 attention backward: errors vs the HMMA backward, then ms at cfg2 shapes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from time_r1_b200 import ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_kernels_gpu import make_mask, rnd

hd = 128
for kind, T, P, C, Hq, Hkv in [("causal", 128, 0, 0, 1, 1), ("causal", 300, 0, 0, 4, 2), ("packed", 531, 203, 41, 14, 2), ("window", 700, 0, 0, 2, 2),
                               ("packed", 5888, 4216, 200, 7, 1)]:
    ld = (Hq + 2 * Hkv) * hd
    qkv = rnd(T, ld, seed=T, scale=0.7)
    q, k, v = qkv[:, :Hq * hd], qkv[:, Hq * hd:(Hq + Hkv) * hd], qkv[:, (Hq + Hkv) * hd:]
    lo, hi, shared = make_mask(kind, T, P, C)
    if kind == "packed" and T == 5888:
        hi[P + 8 * C:] = 0; lo[P + 8 * C:] = 0
    scale = hd ** -0.5
    out, lse = ops.attn_fwd_tc(q, k, v, lo, hi, shared, Hq, Hkv, hd, scale)
    do = rnd(T, Hq * hd, seed=5)
    ref = torch.zeros_like(qkv); got = torch.full_like(qkv, float("nan"))
    qd, kd = Hq * hd, Hkv * hd
    ops.attn_bwd(q, k, v, out, do, lse, lo, hi, shared, Hq, Hkv, hd, scale, ref[:, :qd], ref[:, qd:qd + kd], ref[:, qd + kd:])
    ops.attn_bwd_tc(q, k, v, out, do, lse, lo, hi, shared, Hq, Hkv, hd, scale, got[:, :qd], got[:, qd:qd + kd], got[:, qd + kd:])
    torch.cuda.synchronize()
    res = {"kind": kind, "T": T, "Hq": Hq}
    for name, sl in (("dq", slice(0, qd)), ("dk", slice(qd, qd + kd)), ("dv", slice(qd + kd, None))):
        a, b = got[:, sl].float(), ref[:, sl].float()
        res[name] = {"max_err": float((a - b).abs().nan_to_num(nan=1e9).max()), "ref_absmax": float(b.abs().max()), "nan": int(torch.isnan(a).sum())}
    print(json.dumps(res), flush=True)

T, P, C, Hq, Hkv = 5888, 4216, 200, 28, 4
qkv = rnd(T, (Hq + 2 * Hkv) * hd, seed=1)
qd, kd = Hq * hd, Hkv * hd
q, k, v = qkv[:, :qd], qkv[:, qd:qd + kd], qkv[:, qd + kd:]
lo, hi, shared = make_mask("packed", T, P, C); hi[P + 8 * C:] = 0; lo[P + 8 * C:] = 0
out, lse = ops.attn_fwd_tc(q, k, v, lo, hi, shared, Hq, Hkv, hd, hd ** -0.5)
do = rnd(T, qd, seed=5)
g = torch.empty_like(qkv)
ws = torch.empty(int(__import__("time_r1_b200.lib", fromlist=["x"]).load().t1_attn_bwd_tc_workspace_bytes(T, T, Hq)), device="cuda", dtype=torch.uint8)
for name, fn, kw in (("hmma", ops.attn_bwd, {}), ("tcgen05", ops.attn_bwd_tc, {"ws": ws})):
    for _ in range(2): fn(q, k, v, out, do, lse, lo, hi, shared, Hq, Hkv, hd, hd ** -0.5, g[:, :qd], g[:, qd:qd + kd], g[:, qd + kd:], **kw)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): fn(q, k, v, out, do, lse, lo, hi, shared, Hq, Hkv, hd, hd ** -0.5, g[:, :qd], g[:, qd:qd + kd], g[:, qd + kd:], **kw)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    fl = 2.5 * 4.0 * Hq * hd * (P * (P + 1) / 2 + 8 * sum(P + c + 1 for c in range(C)))
    print(json.dumps({"kernel": name, "ms": round(ms, 3), "tflops_5gemm_equiv": round(fl / ms / 1e9, 1)}), flush=True)
