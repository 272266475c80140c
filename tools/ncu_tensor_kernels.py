"""One launch each of the tensor-bound kernels at cfg2 shapes inside a cudaProfilerStart/Stop window (ncu --set full
--profile-from-start off): the gate_up GEMM [5888 x 37888 x 3584] (K-major) and its wgrad form (MN-major operands),
tcgen05 attention forward and backward over the packed scoring layout."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from time_r1_b200 import ops
from test_kernels_gpu import make_mask, rnd

T, P, C, Hq, Hkv, hd, D, I = 5888, 4216, 200, 28, 4, 128, 3584, 18944
x = rnd(T, D, seed=1); w = rnd(2 * I, D, seed=2, scale=0.02); gu = torch.empty(T, 2 * I, device="cuda", dtype=torch.bfloat16)
dw = torch.empty(2 * I, D, device="cuda", dtype=torch.bfloat16)
qkv = rnd(T, (Hq + 2 * Hkv) * hd, seed=3)
qd, kd = Hq * hd, Hkv * hd
q, k, v = qkv[:, :qd], qkv[:, qd:qd + kd], qkv[:, qd + kd:]
lo, hi, shared = make_mask("packed", T, P, C); hi[P + 8 * C:] = 0; lo[P + 8 * C:] = 0
out = torch.empty(T, qd, device="cuda", dtype=torch.bfloat16); lse = torch.empty(Hq, T, device="cuda")
do = rnd(T, qd, seed=5); g = torch.empty_like(qkv)

act = torch.empty(T, I, device="cuda", dtype=torch.bfloat16)
dx = torch.empty(T, D, device="cuda", dtype=torch.bfloat16)
# ViT: one full-attention block of cfg2 (16 slots x 1024 patches, 16 heads x 80)
from time_r1_b200 import config, indexing
import numpy as np
vcfg = config.PRESETS["qwen2.5-vl-7b"]()
plan = indexing.vision_plan(vcfg, [(16, 32, 32)])
vqkv = rnd(plan.n_patches, 3 * 1280, seed=7)
i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).cuda()
flo, fhi = i32(plan.full_lo), i32(plan.full_hi)
vout = torch.empty(plan.n_patches, 1280, device="cuda", dtype=torch.bfloat16)
xn = rnd(T, D, seed=9); nw = torch.ones(D, device="cuda", dtype=torch.bfloat16); nh = torch.empty_like(xn); nx = torch.empty_like(xn)

def run():
    ops.gemm(x, w, out=gu)                                   # 2-CTA (cta_group::2) GEMM, TMA-store epilogue
    ops.gemm_swiglu(x, w, out=act, gate_up_out=gu)           # the step's gate_up form: SwiGLU epilogue + kept projection
    ops.gemm(gu, w, out=dx, b_t=True)                        # dgrad form (B MN-major), K = 37888
    ops.attn_fwd_tc(vqkv[:, :1280], vqkv[:, 1280:2560], vqkv[:, 2560:], flo, fhi, 0, 16, 16, 80, 80 ** -0.5, out=vout, need_lse=False)
    ops.rmsnorm(xn, nw, 1e-6, delta=x, out=nh, x_out=nx)
    ops.gemm(gu, x, out=dw, a_t=True, b_t=True)
    ops.attn_fwd_tc(q, k, v, lo, hi, shared, Hq, Hkv, hd, hd ** -0.5, out=out, lse=lse)
    ops.attn_bwd_tc(q, k, v, out, do, lse, lo, hi, shared, Hq, Hkv, hd, hd ** -0.5, g[:, :qd], g[:, qd:qd + kd], g[:, qd + kd:])

run(); run(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
