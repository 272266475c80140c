import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from time_r1_b200 import ops
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_kernels_gpu import make_mask, dense_attention, rnd
for kind, T, P, C, Hq, Hkv in [("causal", 128, 0, 0, 1, 1), ("causal", 300, 0, 0, 4, 2), ("packed", 531, 203, 41, 14, 2), ("window", 700, 0, 0, 2, 2), ("packed", 5888, 4216, 200, 4, 1)]:
    hd = 128
    ld = (Hq + 2 * Hkv) * hd
    qkv = rnd(T, ld, seed=T)
    q, k, v = qkv[:, :Hq * hd], qkv[:, Hq * hd:(Hq + Hkv) * hd], qkv[:, (Hq + Hkv) * hd:]
    lo, hi, shared = make_mask(kind, T, P, C)
    if kind == "packed" and T == 5888:
        hi[P + 8 * C:] = 0; lo[P + 8 * C:] = 0
    scale = hd ** -0.5
    ref, lse_ref = ops.attn_fwd(q, k, v, lo, hi, shared, Hq, Hkv, hd, scale)
    out, lse = ops.attn_fwd_tc(q, k, v, lo, hi, shared, Hq, Hkv, hd, scale)
    torch.cuda.synchronize()
    err = (out.float() - ref.float()).abs().max().item()
    m = torch.isfinite(lse_ref)
    lerr = (lse[m] - lse_ref[m]).abs().max().item()
    print(json.dumps({"kind": kind, "T": T, "Hq": Hq, "max_err_vs_hmma": err, "lse_err": lerr, "ref_absmax": ref.float().abs().max().item()}), flush=True)
# timing at cfg2 shapes
T, P, C, Hq, Hkv, hd = 5888, 4216, 200, 28, 4, 128
qkv = rnd(T, (Hq + 2 * Hkv) * hd, seed=1)
q, k, v = qkv[:, :Hq * hd], qkv[:, Hq * hd:(Hq + Hkv) * hd], qkv[:, (Hq + Hkv) * hd:]
lo, hi, shared = make_mask("packed", T, P, C); hi[P + 8 * C:] = 0; lo[P + 8 * C:] = 0
out = torch.empty(T, Hq * hd, device="cuda", dtype=torch.bfloat16); lse = torch.empty(Hq, T, device="cuda")
for name, fn in (("hmma", ops.attn_fwd), ("tcgen05", ops.attn_fwd_tc)):
    for _ in range(3): fn(q, k, v, lo, hi, shared, Hq, Hkv, hd, hd ** -0.5, out=out, lse=lse)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): fn(q, k, v, lo, hi, shared, Hq, Hkv, hd, hd ** -0.5, out=out, lse=lse)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    fl = 4.0 * Hq * hd * (P * (P + 1) / 2 + 8 * sum(P + c + 1 for c in range(C)))
    print(json.dumps({"kernel": name, "ms": round(ms, 3), "tflops": round(fl / ms / 1e9, 1)}), flush=True)
