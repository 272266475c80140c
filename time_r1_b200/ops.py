"""Thin tensor-level wrappers over the libtimer1.so C-ABI (plumbing only: pointer/shape extraction and checks)."""
from __future__ import annotations

import torch

from . import lib
from .lib import call, ptr, stream_ptr


def _chk(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise lib.Timer1Error(f"{name}: tensor must live on a CUDA device (no CPU fallback exists)")
    if t.dtype != dtype:
        raise lib.Timer1Error(f"{name}: expected dtype {dtype}, got {t.dtype}")


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor | None = None, bias: torch.Tensor | None = None,
         accumulate: bool = False, out_dtype=torch.bfloat16) -> torch.Tensor:
    """out[M,N] (+)= a[M,K] @ b[N,K]^T (+ bias[N]) on the tcgen05 kernel. a/b bf16 row-major (row stride free)."""
    _chk(a, torch.bfloat16, "gemm.a")
    _chk(b, torch.bfloat16, "gemm.b")
    M, K = a.shape
    N, K2 = b.shape
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
    if bias is not None:
        assert bias.numel() == N and bias.is_contiguous()
        if bias.dtype == torch.float32:
            flags |= lib.GEMM_BIAS_F32
        elif bias.dtype != torch.bfloat16:
            raise lib.Timer1Error("gemm: bias must be bf16 or fp32")
    call("t1_gemm_bf16_tn", ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out), out.stride(0), M, N, K, ptr(bias),
         flags, stream_ptr())
    return out
