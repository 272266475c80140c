"""Debug trace of one decode_attn_tc CTA (build with T1_EXTRA_NVCC=-DT1_DECODE_DBG)."""
import os, sys, json, ctypes
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv.append("x")
from tools import decode_attn_dbg as D   # runs its timing section too
from time_r1_b200 import lib
from time_r1_b200.lib import call, ptr, stream_ptr
L = lib.load()
fn = L.t1_decode_attn_dbg_read
fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
for S in (1, 8):
    for _ in range(3):
        call("t1_decode_attn_cluster", ptr(D.raw), D.ld, ptr(D.tab), D.maps.ctypes.data, len(D.Ps), D.G, D.Hq, D.Hkv, D.hd, D.cpad, ptr(D.st), ptr(D.out), D.out.stride(0), D.hd ** -0.5, S, stream_ptr())
    torch.cuda.synchronize()
    buf = np.zeros(1024, np.int64)
    assert fn(buf.ctypes.data) == 0
    t0 = buf[900]
    print("S", S, "marks(clk rel start): after_qbuild", buf[901] - t0, "loop_end", buf[902] - t0, "pre_sync1", buf[903] - t0, "post_sync1", buf[904] - t0, "pushed", buf[905] - t0, "post_sync2", buf[906] - t0, "merged", buf[907] - t0)
    n = 12
    for it in range(n):
        r = buf[it * 8: it * 8 + 6] - t0
        print(f"  it {it:2d} tma_issued {r[0]:7d} mma_full {r[1]:7d} mma_pready {r[2]:7d} sm_sfull {r[3]:7d} sm_xch {r[4]:7d} sm_pdone {r[5]:7d}")
