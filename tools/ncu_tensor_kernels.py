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

def run():
    ops.gemm(x, w, out=gu)
    ops.gemm(gu, x, out=dw, a_t=True, b_t=True)
    ops.attn_fwd_tc(q, k, v, lo, hi, shared, Hq, Hkv, hd, hd ** -0.5, out=out, lse=lse)
    ops.attn_bwd_tc(q, k, v, out, do, lse, lo, hi, shared, Hq, Hkv, hd, hd ** -0.5, g[:, :qd], g[:, qd:qd + kd], g[:, qd + kd:])

run(); run(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
