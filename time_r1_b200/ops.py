"""Thin tensor-level wrappers over the libtimer1.so C-ABI (plumbing only: pointer/shape extraction and checks)."""
from __future__ import annotations

import torch

from . import lib
from .lib import call, ptr, stream_ptr


class KernelProfiler:
    """Optional per-launch CUDA-event timing of the GEMM kernels (bench.py roofline). Disabled by default."""

    def __init__(self):
        self.enabled = False
        self.records: list = []   # (kind, flops, bytes, start_event, end_event)

    def start(self):
        self.enabled, self.records = True, []

    def stop(self):
        self.enabled = False

    def begin(self):
        if not self.enabled:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, kind, flops, nbytes, e0):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.records.append((kind, flops, nbytes, e0, e1))

    def summary(self) -> dict:
        torch.cuda.synchronize()
        out: dict = {}
        for kind, fl, nb, e0, e1 in self.records:
            d = out.setdefault(kind, dict(launches=0, flops=0.0, bytes=0.0, ms=0.0))
            d["launches"] += 1
            d["flops"] += fl
            d["bytes"] += nb
            d["ms"] += e0.elapsed_time(e1)
        return out


PROFILER = KernelProfiler()


def _chk(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise lib.Timer1Error(f"{name}: tensor must live on a CUDA device (no CPU fallback exists)")
    if t.dtype != dtype:
        raise lib.Timer1Error(f"{name}: expected dtype {dtype}, got {t.dtype}")


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor | None = None, bias: torch.Tensor | None = None,
         accumulate: bool = False, out_dtype=torch.bfloat16, a_t: bool = False, b_t: bool = False) -> torch.Tensor:
    """out[M,N] (+)= A[M,K] @ B[N,K]^T (+ bias[N]) on the tcgen05 kernel. a/b bf16 row-major (row stride free);
    a_t: `a` holds A^T as [K, M];  b_t: `b` holds B^T as [K, N]  (consumed in place as MN-major UMMA operands)."""
    _chk(a, torch.bfloat16, "gemm.a")
    _chk(b, torch.bfloat16, "gemm.b")
    M, K = (a.shape[1], a.shape[0]) if a_t else a.shape
    N, K2 = (b.shape[1], b.shape[0]) if b_t else b.shape
    assert K == K2, (a.shape, b.shape)
    assert a.stride(1) == 1 and b.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    assert out.shape == (M, N) and out.stride(1) == 1
    flags = 0
    if out.dtype == torch.float32:
        flags |= lib.GEMM_OUT_F32
    elif out.dtype != torch.bfloat16:
        raise lib.Timer1Error("gemm: out must be bf16 or fp32")
    if accumulate:
        flags |= lib.GEMM_ACCUMULATE
    if a_t:
        flags |= lib.GEMM_A_MN
    if b_t:
        flags |= lib.GEMM_B_MN
    if bias is not None:
        assert bias.numel() == N and bias.is_contiguous()
        if bias.dtype == torch.float32:
            flags |= lib.GEMM_BIAS_F32
        elif bias.dtype != torch.bfloat16:
            raise lib.Timer1Error("gemm: bias must be bf16 or fp32")
    e0 = PROFILER.begin()
    call("t1_gemm_bf16_tn", ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out), out.stride(0), M, N, K, ptr(bias),
         flags, stream_ptr())
    PROFILER.end("gemm_bf16_tn", 2.0 * M * N * K, 2.0 * (M * K + N * K) + out.element_size() * M * N, e0)
    return out


BF16 = torch.bfloat16
F32 = torch.float32
I32 = torch.int32


def gemm_swiglu(a, w_gate_up, bias=None, out=None, gate_up_out=None):
    """out[M, I] = silu(a @ Wg^T + bg) * (a @ Wu^T + bu) with w_gate_up = [Wg; Wu] ([2I, K]): gate_up projection + SwiGLU in
    the GEMM epilogue (bit-identical to gemm followed by swiglu); gate_up_out [M, 2I] optionally keeps the projection."""
    M, K = a.shape
    I = w_gate_up.shape[0] // 2
    assert a.stride(1) == 1 and w_gate_up.stride(1) == 1 and w_gate_up.shape[1] == K and I % 8 == 0
    if out is None:
        out = torch.empty((M, I), device=a.device, dtype=BF16)
    assert out.shape == (M, I) and out.dtype == BF16 and out.stride(1) == 1
    if gate_up_out is not None:
        assert gate_up_out.shape == (M, 2 * I) and gate_up_out.dtype == BF16 and gate_up_out.stride(1) == 1
    if bias is not None:
        assert bias.dtype == BF16 and bias.numel() == 2 * I and bias.is_contiguous()
    e0 = PROFILER.begin()
    call("t1_gemm_swiglu", ptr(a), a.stride(0), ptr(w_gate_up), w_gate_up.stride(0), ptr(bias), ptr(out), out.stride(0),
         ptr(gate_up_out), gate_up_out.stride(0) if gate_up_out is not None else 0, M, I, K, stream_ptr())
    PROFILER.end("gemm_bf16_tn", 4.0 * M * I * K, 2.0 * (M * K + 2 * I * K) + 2.0 * M * I, e0)
    return out


def skinny_gemm(x, w, out=None, bias=None, accumulate=False, out_dtype=BF16):
    """out[M<=64, N] (+)= x[M,K] @ w[N,K]^T (+bias): weight-streaming decode GEMM."""
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=out_dtype)
    call("t1_skinny_gemm", ptr(x), x.stride(0), ptr(w), w.stride(0), ptr(bias), ptr(out), out.stride(0), M, N, K,
         1 if out.dtype == F32 else 0, 1 if accumulate else 0, stream_ptr())
    return out


def stream_gemm_workspace(max_m: int, device) -> torch.Tensor:
    """Zero-filled workspace for stream_gemm launches of up to max_m rows on one stream."""
    n = int(lib.load().t1_stream_gemm_workspace_bytes(int(max_m)))
    return torch.zeros(n, device=device, dtype=torch.uint8)


def stream_gemm(x, w, workspace, out=None, bias=None, out_dtype=BF16):
    """out[M<=64, N] = x[M,K] @ w[N,K]^T (+bias) on the stream-K tcgen05 weight-streaming kernel."""
    M, K = x.shape
    N = w.shape[0]
    assert x.stride(1) == 1 and w.stride(1) == 1 and w.shape[1] == K
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=out_dtype)
    assert out.shape == (M, N) and out.stride(1) == 1
    call("t1_stream_gemm", ptr(x), x.stride(0), ptr(w), w.stride(0), ptr(bias), ptr(out), out.stride(0), M, N, K,
         1 if out.dtype == F32 else 0, ptr(workspace), workspace.numel(), stream_ptr())
    return out


def stream_gemm_swiglu(x, w_gate_up, workspace, out=None, bias=None):
    """out[M<=64, I] = silu(x @ Wg^T) * (x @ Wu^T) with w_gate_up = [Wg; Wu] ([2I, K]): gate_up projection + SwiGLU in
    one weight stream (same per-element arithmetic as stream_gemm followed by swiglu)."""
    M, K = x.shape
    I = w_gate_up.shape[0] // 2
    assert x.stride(1) == 1 and w_gate_up.stride(1) == 1 and w_gate_up.shape[1] == K
    if out is None:
        out = torch.empty((M, I), device=x.device, dtype=BF16)
    assert out.shape == (M, I) and out.stride(1) == 1 and out.dtype == BF16
    call("t1_stream_gemm_swiglu", ptr(x), x.stride(0), ptr(w_gate_up), w_gate_up.stride(0), ptr(bias), ptr(out), out.stride(0),
         M, I, K, ptr(workspace), workspace.numel(), stream_ptr())
    return out


def stream_gemm_resnorm(x, w, workspace, resid, norm_w, eps, h_out, out=None):
    """out = x @ w^T;  resid += out (in place);  h_out = rmsnorm(resid, eps) * norm_w — projection, residual add and the
    next RMSNorm of a decode layer in one launch."""
    M, K = x.shape
    N = w.shape[0]
    assert x.stride(1) == 1 and w.stride(1) == 1 and w.shape[1] == K
    assert resid.shape == (M, N) and resid.is_contiguous() and h_out.shape == (M, N) and h_out.is_contiguous()
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=BF16)
    assert out.dtype == BF16 and out.stride(1) == 1
    call("t1_stream_gemm_resnorm", ptr(x), x.stride(0), ptr(w), w.stride(0), ptr(out), out.stride(0), M, N, K, ptr(resid),
         ptr(norm_w), float(eps), ptr(h_out), ptr(workspace), workspace.numel(), stream_ptr())
    return out


def stream_gemm_fold(x, w, workspace, resid, norm_w, z_out, parts_out):
    """resid += bf16(x @ w^T) in place; z_out = bf16(norm_w * resid); parts_out[N/128, MP] = per-row partial sums of
    resid^2 — the producer half of an RMSNorm folded into the GEMMs around it (see stream_gemm_scaled)."""
    M, K = x.shape
    N = w.shape[0]
    MP = 16 if M <= 16 else (32 if M <= 32 else 64)
    assert x.stride(1) == 1 and w.stride(1) == 1 and w.shape[1] == K and N % 128 == 0
    assert resid.shape == (M, N) and resid.is_contiguous() and z_out.shape == (M, N) and z_out.is_contiguous()
    assert parts_out.dtype == F32 and parts_out.is_contiguous() and parts_out.numel() >= (N // 128) * MP
    call("t1_stream_gemm_fold", ptr(x), x.stride(0), ptr(w), w.stride(0), M, N, K, ptr(resid), ptr(norm_w), ptr(z_out),
         ptr(parts_out), ptr(workspace), workspace.numel(), stream_ptr())
    return z_out


def stream_gemm_scaled(z, w, workspace, parts, n_parts, eps, out=None, bias=None, swiglu=False):
    """out = rmsnorm-row-factor * (z @ w^T) (+bias) [then SwiGLU]: the consumer half of a folded RMSNorm; z / parts come
    from stream_gemm_fold (n_parts = hidden / 128)."""
    M, K = z.shape
    N = w.shape[0]
    cols = N // 2 if swiglu else N
    assert z.stride(1) == 1 and w.stride(1) == 1 and w.shape[1] == K and parts.dtype == F32
    if out is None:
        out = torch.empty((M, cols), device=z.device, dtype=BF16)
    assert out.shape == (M, cols) and out.stride(1) == 1 and out.dtype == BF16
    call("t1_stream_gemm_scaled", ptr(z), z.stride(0), ptr(w), w.stride(0), ptr(bias), ptr(out), out.stride(0), M, N, K,
         1 if swiglu else 0, ptr(parts), int(n_parts), float(eps), ptr(workspace), workspace.numel(), stream_ptr())
    return out


def stream_gemm_fold_pair(x, w_a, workspace, resid, norm_w, z, parts, w_b, eps, out=None, bias=None, swiglu=False):
    """stream_gemm_fold(x, w_a -> resid, z, parts) + stream_gemm_scaled(z, w_b, parts -> out) in ONE launch (bit-identical to
    the two): the boundary between the GEMMs is a device-side arrival counter and the weight stream runs through it."""
    M, K = x.shape
    Na, Nb = w_a.shape[0], w_b.shape[0]
    MP = 16 if M <= 16 else (32 if M <= 32 else 64)
    cols = Nb // 2 if swiglu else Nb
    assert x.stride(1) == 1 and w_a.stride(1) == 1 and w_b.stride(1) == 1 and w_a.shape[1] == K and w_b.shape[1] == Na and Na % 128 == 0
    assert resid.shape == (M, Na) and resid.is_contiguous() and z.shape == (M, Na) and z.is_contiguous()
    assert parts.dtype == F32 and parts.is_contiguous() and parts.numel() >= (Na // 128) * MP
    if out is None:
        out = torch.empty((M, cols), device=x.device, dtype=BF16)
    assert out.shape == (M, cols) and out.stride(1) == 1 and out.dtype == BF16
    call("t1_stream_gemm_fold_pair", ptr(x), x.stride(0), ptr(w_a), w_a.stride(0), M, Na, K, ptr(resid), ptr(norm_w), ptr(z), ptr(parts),
         ptr(w_b), w_b.stride(0), ptr(bias), ptr(out), out.stride(0), Nb, 1 if swiglu else 0, float(eps), ptr(workspace),
         workspace.numel(), stream_ptr())
    return out


def linear(x, w, out=None, bias=None, accumulate=False, out_dtype=BF16):
    """Dispatch on the row count: <= 64 rows stream the weights (HBM-bound), otherwise the tcgen05 tile kernel."""
    if x.shape[0] <= 64 and w.shape[0] % 16 == 0 and x.shape[1] % 32 == 0 and (bias is None or bias.dtype == BF16):
        return skinny_gemm(x, w, out=out, bias=bias, accumulate=accumulate, out_dtype=out_dtype)
    return gemm(x, w, out=out, bias=bias, accumulate=accumulate, out_dtype=out_dtype)


def rmsnorm(x, w, eps, delta=None, out=None, x_out=None):
    """y = rmsnorm(x + delta) * w.  With delta, the summed residual stream is written to x_out (default: x in place)."""
    T, D = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    if delta is not None and x_out is None:
        x_out = x
    call("t1_rmsnorm", ptr(x), ptr(delta), ptr(x_out), ptr(w), ptr(out), T, D, float(eps), stream_ptr())
    return out


def layernorm(x, w, b, eps, delta=None, out=None, x_out=None):
    T, D = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    if delta is not None and x_out is None:
        x_out = x
    call("t1_layernorm", ptr(x), ptr(delta), ptr(x_out), ptr(w), ptr(b), ptr(out), T, D, float(eps), stream_ptr())
    return out


def gather_rows_f32(src, idx, out=None):
    """out_bf16[i] = bf16(src_f32[idx[i]]) — pixel rows into ViT window order."""
    n = idx.numel() if idx is not None else src.shape[0]
    width = src.shape[1]
    if out is None:
        out = torch.empty((n, width), device=src.device, dtype=BF16)
    call("t1_gather_rows_f32", ptr(src), src.stride(0), ptr(idx), ptr(out), out.stride(0), n, width, stream_ptr())
    return out


def rmsnorm_bwd(dy, x, w, eps, dw_f32, dres=None, out=None):
    """dx = (dres +) d rmsnorm(x)·w / dx ; dw_f32 (+)= column reduction (fp32 atomics)."""
    T, D = x.shape
    if out is None:
        out = torch.empty_like(x)
    call("t1_rmsnorm_bwd", ptr(dy), ptr(x), ptr(w), ptr(dres), ptr(out), ptr(dw_f32), T, D, float(eps), stream_ptr())
    return out


def layernorm_bwd(dy, x, w, eps, dw_f32, db_f32, need_dx=True, out=None):
    """LayerNorm backward (Qwen2-VL merger and ViT blocks): dw_f32 / db_f32 accumulate; returns dx (or None when need_dx is False)."""
    T, D = x.shape
    _chk(dy, BF16, "dy"); _chk(x, BF16, "x")
    if need_dx and out is None:
        out = torch.empty_like(x)
    if not need_dx:
        out = None
    call("t1_layernorm_bwd", ptr(dy), ptr(x), ptr(w), ptr(out), ptr(dw_f32), ptr(db_f32), T, D, float(eps), stream_ptr())
    return out


def rope_table(pos, axis_id, inv_freq):
    """pos [n_axes, T] int32 -> (cos, sin) [T, half] fp32 with channel i using axis axis_id[i]."""
    T = pos.shape[1]
    half = inv_freq.numel()
    cos = torch.empty((T, half), device=pos.device, dtype=F32)
    sin = torch.empty((T, half), device=pos.device, dtype=F32)
    call("t1_rope_table", ptr(pos), T, ptr(axis_id), ptr(inv_freq), half, ptr(cos), ptr(sin), stream_ptr())
    return cos, sin


def rope_apply(x, col0, n_heads, head_dim, cos, sin, inverse=False, out=None):
    """Rotate n_heads heads starting at column col0 of x [T, ld] in place (optionally copying them to out)."""
    T = x.shape[0]
    call("t1_rope_apply", ptr(x), x.stride(0), col0, n_heads, head_dim, ptr(cos), ptr(sin), 1 if inverse else 0, T,
         ptr(out), 0 if out is None else out.stride(0), stream_ptr())


def copy_cols(src, dst, width):
    call("t1_copy_cols", ptr(src), src.stride(0), ptr(dst), dst.stride(0), src.shape[0], width, stream_ptr())


def swiglu(gate_up, out=None):
    T, I2 = gate_up.shape
    assert gate_up.is_contiguous()
    if out is None:
        out = torch.empty((T, I2 // 2), device=gate_up.device, dtype=BF16)
    call("t1_swiglu", ptr(gate_up), ptr(out), T, I2 // 2, stream_ptr())
    return out


def swiglu_bwd(dact, gate_up, out=None):
    T, I2 = gate_up.shape
    if out is None:
        out = torch.empty_like(gate_up)
    call("t1_swiglu_bwd", ptr(dact), ptr(gate_up), ptr(out), T, I2 // 2, stream_ptr())
    return out


def act(x, mode, out=None):
    """mode 0: exact GELU; 1: QuickGELU."""
    assert x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    call("t1_act", ptr(x), ptr(out), x.numel(), mode, stream_ptr())
    return out


def act_bwd(dy, x, mode, out=None):
    if out is None:
        out = torch.empty_like(x)
    call("t1_act_bwd", ptr(dy), ptr(x), ptr(out), x.numel(), mode, stream_ptr())
    return out


def gather_rows(src, idx, out=None, width=None):
    n = idx.numel() if idx is not None else src.shape[0]
    width = width or src.shape[1]
    if out is None:
        out = torch.empty((n, width), device=src.device, dtype=BF16)
    call("t1_gather_rows", ptr(src), src.stride(0), ptr(idx), ptr(out), out.stride(0), n, width, stream_ptr())
    return out


def scatter_rows(src, idx, dst, accumulate=False, width=None):
    n = idx.numel() if idx is not None else src.shape[0]
    width = width or src.shape[1]
    call("t1_scatter_rows", ptr(src), src.stride(0), ptr(idx), ptr(dst), dst.stride(0), n, width,
         1 if accumulate else 0, stream_ptr())
    return dst


def transpose(x, out=None):
    R, C = x.shape
    if out is None:
        out = torch.empty((C, R), device=x.device, dtype=BF16)
    call("t1_transpose_bf16", ptr(x), x.stride(0), ptr(out), out.stride(0), R, C, stream_ptr())
    return out


def colsum(x, out_f32):
    R, C = x.shape
    call("t1_colsum", ptr(x), x.stride(0), R, C, ptr(out_f32), stream_ptr())
    return out_f32


def f32_to_bf16(src, dst, accumulate=False):
    call("t1_f32_to_bf16", ptr(src), ptr(dst), src.numel(), 1 if accumulate else 0, stream_ptr())
    return dst


def add_(y, x):
    call("t1_add_bf16", ptr(y), ptr(x), y.numel(), stream_ptr())
    return y


def attn_fwd(q, k, v, row_lo, row_hi, shared_len, Hq, Hkv, head_dim, scale, out=None, lse=None, need_lse=True):
    """q: [T, >=Hq*hd] view, k/v: [Tk, >=Hkv*hd] views (row strides free). Returns (out [T, Hq*hd], lse2 [Hq, T])."""
    T, Tk = q.shape[0], k.shape[0]
    if out is None:
        out = torch.empty((T, Hq * head_dim), device=q.device, dtype=BF16)
    if lse is None and need_lse:
        lse = torch.empty((Hq, T), device=q.device, dtype=F32)
    call("t1_attn_fwd", ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(out), out.stride(0),
         ptr(lse), ptr(row_lo), ptr(row_hi), int(shared_len), T, Tk, Hq, Hkv, head_dim, float(scale), stream_ptr())
    return out, lse


def attn_fwd_tc(q, k, v, row_lo, row_hi, shared_len, Hq, Hkv, head_dim, scale, out=None, lse=None, need_lse=True,
                q_row0=0):
    """tcgen05/TMEM flash-attention forward (head_dim 128); same contract as attn_fwd.  With q_row0 > 0 only the query
    rows [q_row0, Tk) are processed and q / out hold just those rows (lo / hi / lse keep absolute row indices)."""
    Tk = k.shape[0]
    T = q.shape[0] + q_row0
    if out is None:
        out = torch.empty((q.shape[0], Hq * head_dim), device=q.device, dtype=BF16)
    if lse is None and need_lse:
        lse = torch.empty((Hq, T), device=q.device, dtype=F32)
    call("t1_attn_fwd_tc", ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(out), out.stride(0),
         ptr(lse), lse.stride(0) if lse is not None else 0, ptr(row_lo), ptr(row_hi), int(shared_len), T, Tk, Hq, Hkv, head_dim,
         float(scale), int(q_row0), stream_ptr())
    return out, lse


def attn_fwd_tc_win(q, k, v, row_lo, row_hi, H, head_dim, scale, out=None):
    """Persistent tcgen05 attention for block-local masks (each row's [lo, hi) inside its own 128-row tile: the ViT's windows)."""
    T = q.shape[0]
    if out is None:
        out = torch.empty((T, H * head_dim), device=q.device, dtype=BF16)
    call("t1_attn_fwd_tc_win", ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(out), out.stride(0), ptr(row_lo),
         ptr(row_hi), T, H, head_dim, float(scale), stream_ptr())
    return out


def attn_bwd(q, k, v, o, do, lse, row_lo, row_hi, shared_len, Hq, Hkv, head_dim, scale, dq, dk, dv):
    T, Tk = q.shape[0], k.shape[0]
    delta = torch.empty((Hq, T), device=q.device, dtype=F32)
    summ = torch.empty(((T + 63) // 64) * 4, device=q.device, dtype=I32)
    call("t1_attn_bwd", ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(o), o.stride(0), ptr(do),
         do.stride(0), ptr(lse), ptr(delta), ptr(summ), ptr(dq), dq.stride(0), ptr(dk), dk.stride(0), ptr(dv),
         dv.stride(0), ptr(row_lo), ptr(row_hi), int(shared_len), T, Tk, Hq, Hkv, head_dim, float(scale), stream_ptr())
    return dq, dk, dv


def attn_bwd_tc(q, k, v, o, do, lse, row_lo, row_hi, shared_len, Hq, Hkv, head_dim, scale, dq, dk, dv, ws=None):
    """tcgen05/TMEM flash-attention backward (head_dim 128); same contract as attn_bwd."""
    T, Tk = q.shape[0], k.shape[0]
    need = int(lib.load().t1_attn_bwd_tc_workspace_bytes(T, Tk, Hq))
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, device=q.device, dtype=torch.uint8)
    call("t1_attn_bwd_tc", ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(o), o.stride(0), ptr(do),
         do.stride(0), ptr(lse), ptr(ws), ws.numel(), ptr(dq), dq.stride(0), ptr(dk), dk.stride(0), ptr(dv), dv.stride(0),
         ptr(row_lo), ptr(row_hi), int(shared_len), T, Tk, Hq, Hkv, head_dim, float(scale), stream_ptr())
    return dq, dk, dv


OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def video_patchify(frames, patch=14, tpatch=2, merge=2, mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD, rescale_factor=1.0 / 255.0,
                   perm_rows=None, want_f32=True, want_bf16=False):
    """frames [T, 3, H, W] uint8 or fp32 (0..255) on the device -> (pixel_values_videos [N, 1176] fp32 in HF order or None,
    bf16 [N, 1176] gathered through perm_rows or None, (grid_t, grid_h, grid_w))."""
    import ctypes
    import numpy as np

    assert frames.is_cuda and frames.dim() == 4 and frames.shape[1] == 3 and frames.is_contiguous()
    assert frames.dtype in (torch.uint8, F32)
    T, _, H, W = frames.shape
    gt, gh, gw = (T + tpatch - 1) // tpatch, H // patch, W // patch
    N, row = gt * gh * gw, 3 * tpatch * patch * patch
    # fused rescale-and-normalise constants exactly as the HF processor forms them: fp32 tensor * python float
    m255 = (torch.tensor(mean, dtype=F32) * (1.0 / rescale_factor)).numpy().astype(np.float32)
    s255 = (torch.tensor(std, dtype=F32) * (1.0 / rescale_factor)).numpy().astype(np.float32)
    o32 = torch.empty((N, row), device=frames.device, dtype=F32) if want_f32 else None
    o16 = torch.empty((N, row), device=frames.device, dtype=BF16) if want_bf16 else None
    call("t1_video_patchify", ptr(frames), 1 if frames.dtype == torch.uint8 else 0, T, H, W, patch, tpatch, merge,
         m255.ctypes.data_as(ctypes.c_void_p), s255.ctypes.data_as(ctypes.c_void_p), ptr(perm_rows), ptr(o32), ptr(o16), stream_ptr())
    return o32, o16, (gt, gh, gw)


def kv_cache_tensor_maps(caches, Hkv, head_dim):
    """Host table (numpy [n, 128] uint8) of 128-byte TMA tensor maps (box 128 rows x 64 columns, SWIZZLE_128B) over a list of
    [rows, Hkv*head_dim] bf16 cache tensors, in list order — the `maps` argument of t1_decode_attn_cluster."""
    import ctypes
    import numpy as np

    buf = np.zeros((len(caches), 128), dtype=np.uint8)
    for i, c in enumerate(caches):
        assert c.dtype == BF16 and c.stride(1) == 1 and c.shape[1] == Hkv * head_dim
        call("t1_encode_tmap_bf16", buf[i].ctypes.data_as(ctypes.c_void_p), ptr(c), c.shape[0], c.shape[1], c.stride(0), 128, 64)
    return buf


def decode_attn_wide_workspace(n_groups, Hkv, device, splits=0):
    """(splits, zeroed uint8 workspace) for t1_decode_attn_wide on `device` (plan: t1_decode_attn_wide_plan)."""
    import ctypes

    s, nb = ctypes.c_int(0), ctypes.c_longlong(0)
    call("t1_decode_attn_wide_plan", int(n_groups), int(Hkv), int(splits), ctypes.byref(s), ctypes.byref(nb))
    return s.value, torch.zeros(nb.value, device=device, dtype=torch.uint8)


def logprob_entropy(logits, target):
    R, V = logits.shape
    lp = torch.empty(R, device=logits.device, dtype=F32)
    ent = torch.empty(R, device=logits.device, dtype=F32)
    lse = torch.empty(R, device=logits.device, dtype=F32)
    call("t1_logprob_entropy", ptr(logits), logits.stride(0), R, V, ptr(target), ptr(lp), ptr(ent), ptr(lse), stream_ptr())
    return lp, ent, lse


def head_logprob_entropy(h, w, target):
    """log p(target), entropy and log-sum-exp per row of softmax(h @ w^T) WITHOUT the logits tensor: the tcgen05 GEMM keeps per
    256-column tile and row (max, sum exp, sum exp * x, target logit) and t1_lse_merge folds the tiles."""
    M, K = h.shape
    V = w.shape[0]
    assert h.dtype == BF16 and w.dtype == BF16 and h.stride(1) == 1 and w.stride(1) == 1 and w.shape[1] == K
    assert target.dtype == I32 and target.numel() == M and target.is_contiguous()
    n_tiles = (V + 255) // 256
    parts = torch.empty((n_tiles, M, 4), device=h.device, dtype=F32)
    e0 = PROFILER.begin()
    call("t1_gemm_lse_parts", ptr(h), h.stride(0), ptr(w), w.stride(0), M, V, K, ptr(target), ptr(parts), stream_ptr())
    PROFILER.end("gemm_bf16_tn", 2.0 * M * V * K, 2.0 * (M * K + V * K) + 16.0 * n_tiles * M, e0)
    lp = torch.empty(M, device=h.device, dtype=F32)
    ent = torch.empty(M, device=h.device, dtype=F32)
    lse = torch.empty(M, device=h.device, dtype=F32)
    call("t1_lse_merge", ptr(parts), n_tiles, M, ptr(lp), ptr(ent), ptr(lse), stream_ptr())
    return lp, ent, lse


def head_dlogits(h, w, target, lse, dlogp, out=None):
    """bf16 dlogits = dlogp * (onehot(target) - softmax(h @ w^T)) straight from the GEMM accumulator (no fp32 logits)."""
    M, K = h.shape
    V = w.shape[0]
    if out is None:
        out = torch.empty((M, V), device=h.device, dtype=BF16)
    assert out.shape == (M, V) and out.dtype == BF16 and out.stride(1) == 1
    assert target.dtype == I32 and lse.dtype == F32 and dlogp.dtype == F32
    e0 = PROFILER.begin()
    call("t1_gemm_dlogits", ptr(h), h.stride(0), ptr(w), w.stride(0), ptr(out), out.stride(0), M, V, K, ptr(target), ptr(lse),
         ptr(dlogp), stream_ptr())
    PROFILER.end("gemm_bf16_tn", 2.0 * M * V * K, 2.0 * (M * K + V * K) + 2.0 * M * V, e0)
    return out


def logits_backward(logits, target, lse, dlogp, out=None):
    R, V = logits.shape
    if out is None:
        out = torch.empty((R, V), device=logits.device, dtype=BF16)
    call("t1_logits_backward", ptr(logits), logits.stride(0), R, V, ptr(target), ptr(lse), ptr(dlogp), ptr(out),
         out.stride(0), stream_ptr())
    return out


def iou_reward(mode, pred_s, pred_e, valid, gt_s, gt_e, dur):
    """All inputs float64 device tensors (valid: uint8). Returns (reward64, reward32)."""
    n = pred_s.numel()
    o64 = torch.empty(n, device=pred_s.device, dtype=torch.float64)
    o32 = torch.empty(n, device=pred_s.device, dtype=F32)
    call("t1_iou_reward", int(mode), ptr(pred_s), ptr(pred_e), ptr(valid), ptr(gt_s), ptr(gt_e), ptr(dur), ptr(o64),
         ptr(o32), n, stream_ptr())
    return o64, o32


def detokenize(ids, table, row_tokens=None, skip_special_tokens=True):
    """ids [n, C] int32 (device) -> (bytes [n, cap] uint8, byte_len [n] int32): the byte string batch_decode builds
    (timer1_trainer.py:642-644).  `table`: detok.DetokTable on the same device."""
    n, C = ids.shape
    _chk(ids, I32, "ids")
    cap = max(16, -(-C * max(1, table.max_token_bytes) // 16) * 16)
    out = torch.empty((n, cap), device=ids.device, dtype=torch.uint8)
    out_len = torch.empty(n, device=ids.device, dtype=I32)
    call("t1_detokenize", ptr(ids), n, C, ptr(row_tokens), ptr(table.offsets), ptr(table.data),
         ptr(table.skip) if skip_special_tokens else None, table.vocab_size, table.max_token_bytes, ptr(out), cap, ptr(out_len),
         stream_ptr())
    return out, out_len


def parse_answers(text, text_len, mode=0):
    """text [n, cap] uint8 + text_len [n] int32 -> (start [n] f64, end [n] f64, valid [n] u8, format [n] u8, format [n] f32);
    mode 0 = training rewards (main.py:122-142, 234-239), 1 / 2 = the filtering epochs' extract_answer (+ force fallback)."""
    n, cap = text.shape
    dev = text.device
    s = torch.empty(n, device=dev, dtype=torch.float64)
    e = torch.empty(n, device=dev, dtype=torch.float64)
    valid = torch.empty(n, device=dev, dtype=torch.uint8)
    fmt = torch.empty(n, device=dev, dtype=torch.uint8)
    fmt32 = torch.empty(n, device=dev, dtype=F32)
    call("t1_parse_answers", ptr(text), ptr(text_len), n, cap, int(mode), ptr(s), ptr(e), ptr(valid), ptr(fmt), ptr(fmt32),
         stream_ptr())
    return s, e, valid, fmt, fmt32


def group_advantage(rewards_per_func, G, eps=1e-4):
    n, nf = rewards_per_func.shape
    dev = rewards_per_func.device
    rewards = torch.empty(n, device=dev, dtype=F32)
    adv = torch.empty(n, device=dev, dtype=F32)
    mean = torch.empty(n // G, device=dev, dtype=F32)
    std = torch.empty(n // G, device=dev, dtype=F32)
    call("t1_group_advantage", ptr(rewards_per_func), n, nf, G, float(eps), ptr(rewards), ptr(adv), ptr(mean), ptr(std),
         stream_ptr())
    return rewards, adv, mean, std


def grpo_loss(logp, ref_logp, entropy, adv, mask, beta, eps_lo, eps_hi, use_grpo, loss_scale=1.0, old_logp=None):
    rows, C = logp.shape
    dev = logp.device
    loss = torch.empty(1, device=dev, dtype=F32)
    dlogp = torch.empty((rows, C), device=dev, dtype=F32)
    stats = torch.empty(8, device=dev, dtype=F32)
    call("t1_grpo_loss", ptr(logp), ptr(old_logp), ptr(ref_logp), ptr(entropy), ptr(adv), ptr(mask), rows, C,
         float(beta), float(eps_lo), float(eps_hi), 1 if use_grpo else 0, float(loss_scale), ptr(loss), ptr(dlogp),
         ptr(stats), stream_ptr())
    return loss, dlogp, stats


def eos_mask(ids, eos_id):
    rows, C = ids.shape
    mask = torch.empty((rows, C), device=ids.device, dtype=I32)
    lengths = torch.empty(rows, device=ids.device, dtype=I32)
    call("t1_eos_mask", ptr(ids), rows, C, int(eos_id), ptr(mask), ptr(lengths), stream_ptr())
    return mask, lengths


def sumsq(g, out_accum):
    call("t1_sumsq_bf16", ptr(g), g.numel(), ptr(out_accum), stream_ptr())
    return out_accum


def adamw(master, m, v, grad, param, lr, beta1, beta2, eps, wd, step, gnorm_sq=None, max_norm=0.0, grad_scale=1.0):
    call("t1_adamw", ptr(master), ptr(m), ptr(v), ptr(grad), ptr(param), master.numel(), float(lr), float(beta1),
         float(beta2), float(eps), float(wd), int(step), ptr(gnorm_sq), float(max_norm), float(grad_scale), stream_ptr())


def bf16_to_f32(src, dst):
    call("t1_bf16_to_f32", ptr(src), ptr(dst), src.numel(), stream_ptr())
    return dst
