"""ctypes binding of libtimer1.so — the C-ABI declared in include/timer1.h.

There is no fallback: if the shared library is missing or fails to load, importing a compute entry point raises.
Tensors cross the boundary as raw device pointers (`tensor.data_ptr()`), sizes as plain ints and the CUDA stream
as a `cudaStream_t` handle; no torch types appear in any signature.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtimer1.so")

_lib = None

vp = C.c_void_p
i32 = C.c_int
i64 = C.c_longlong
f32 = C.c_float
u64 = C.c_ulonglong

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGNATURES: dict[str, list] = {
    "t1_version": [],
    "t1_device_info": [C.POINTER(C.c_int)],
    "t1_count_launch": [],
    "t1_launch_count": [i32],
    # gemm_tcgen05.cu
    "t1_gemm_bf16_tn": [vp, i64, vp, i64, vp, i64, i32, i32, i32, vp, i32, vp],
    "t1_gemm_swiglu": [vp, i64, vp, i64, vp, vp, i64, vp, i64, i32, i32, i32, vp],
    # grpo.cu
    "t1_iou_reward": [i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp],
    "t1_group_advantage": [vp, i32, i32, i32, f32, vp, vp, vp, vp, vp],
    "t1_grpo_loss": [vp, vp, vp, vp, vp, vp, i32, i32, f32, f32, f32, i32, f32, vp, vp, vp, vp],
    "t1_logprob_entropy": [vp, i64, i32, i32, vp, vp, vp, vp, vp],
    "t1_logits_backward": [vp, i64, i32, i32, vp, vp, vp, vp, i64, vp],
    "t1_gemm_lse_parts": [vp, i64, vp, i64, i32, i32, i32, vp, vp, vp],
    "t1_lse_merge": [vp, i32, i32, vp, vp, vp, vp],
    "t1_gemm_dlogits": [vp, i64, vp, i64, vp, i64, i32, i32, i32, vp, vp, vp, vp],
    "t1_eos_mask": [vp, i32, i32, i32, vp, vp, vp],
    # answer_scan.cu
    "t1_detokenize": [vp, i32, i32, vp, vp, vp, vp, i32, i32, vp, i32, vp, vp],
    "t1_parse_answers": [vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp],
    # elementwise.cu
    "t1_rmsnorm": [vp, vp, vp, vp, vp, i32, i32, f32, vp],
    "t1_layernorm": [vp, vp, vp, vp, vp, vp, i32, i32, f32, vp],
    "t1_rmsnorm_bwd": [vp, vp, vp, vp, vp, vp, i32, i32, f32, vp],
    "t1_layernorm_bwd": [vp, vp, vp, vp, vp, vp, i32, i32, f32, vp],
    "t1_rope_table": [vp, i32, vp, vp, i32, vp, vp, vp],
    "t1_rope_apply": [vp, i64, i32, i32, i32, vp, vp, i32, i32, vp, i64, vp],
    "t1_copy_cols": [vp, i64, vp, i64, i32, i32, vp],
    "t1_swiglu": [vp, vp, i64, i32, vp],
    "t1_swiglu_bwd": [vp, vp, vp, i64, i32, vp],
    "t1_act": [vp, vp, i64, i32, vp],
    "t1_act_bwd": [vp, vp, vp, i64, i32, vp],
    "t1_gather_rows": [vp, i64, vp, vp, i64, i32, i32, vp],
    "t1_gather_rows_f32": [vp, i64, vp, vp, i64, i32, i32, vp],
    "t1_scatter_rows": [vp, i64, vp, vp, i64, i32, i32, i32, vp],
    "t1_transpose_bf16": [vp, i64, vp, i64, i32, i32, vp],
    "t1_colsum": [vp, i64, i32, i32, vp, vp],
    "t1_f32_to_bf16": [vp, vp, i64, i32, vp],
    "t1_add_bf16": [vp, vp, i64, vp],
    # attention.cu
    "t1_attn_fwd": [vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, vp],
    "t1_attn_fwd_tc": [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, i32, i32, i32, i32, i32, i32, f32, i32, vp],
    "t1_attn_fwd_tc_win": [vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, i32, i32, i32, f32, vp],
    "t1_decode_attn_cluster": [vp, i64, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, i64, f32, i32, vp],
    "t1_decode_attn_wide_plan": [i32, i32, i32, vp, vp],
    "t1_decode_attn_wide": [vp, i64, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, i64, f32, i32, vp, i64, vp],
    "t1_encode_tmap_bf16": [vp, vp, i64, i64, i64, i32, i32],
    "t1_video_patchify": [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp],
    "t1_attn_bwd_tc_workspace_bytes": [i32, i32, i32],
    "t1_attn_bwd_tc": [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp,
                       i32, i32, i32, i32, i32, i32, f32, vp],
    "t1_attn_bwd": [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, vp, i64, vp, i64, vp, i64, vp, vp,
                    i32, i32, i32, i32, i32, i32, f32, vp],
    # decode.cu
    "t1_skinny_gemm": [vp, i64, vp, i64, vp, vp, i64, i32, i32, i32, i32, i32, vp],
    "t1_stream_gemm": [vp, i64, vp, i64, vp, vp, i64, i32, i32, i32, i32, vp, i64, vp],
    "t1_gemm_max_ctas": [i32],
    "t1_stream_gemm_dynamic": [i32],
    "t1_stream_gemm_max_ctas": [i32],
    "t1_stream_gemm_resnorm": [vp, i64, vp, i64, vp, i64, i32, i32, i32, vp, vp, f32, vp, vp, i64, vp],
    "t1_stream_gemm_swiglu": [vp, i64, vp, i64, vp, vp, i64, i32, i32, i32, vp, i64, vp],
    "t1_stream_gemm_fold": [vp, i64, vp, i64, i32, i32, i32, vp, vp, vp, vp, vp, i64, vp],
    "t1_stream_gemm_scaled": [vp, i64, vp, i64, vp, vp, i64, i32, i32, i32, i32, vp, i32, f32, vp, i64, vp],
    "t1_stream_gemm_fold_pair": [vp, i64, vp, i64, i32, i32, i32, vp, vp, vp, vp, vp, i64, vp, vp, i64, i32, i32, f32, vp, i64, vp],
    "t1_stream_gemm_prefetch_next": [vp, i64, i32, i32, i32, i32, i32, i64],
    "t1_stream_gemm_workspace_bytes": [i32],
    "t1_set_pdl": [i32],
    "t1_stream_gemm_tune": [i32, i32],
    "t1_stream_gemm_debug": [vp, i32],
    "t1_decode_rope_append_groups": [vp, i64, i32, i32, i32, i32, i32, vp, vp, i32, vp],
    "t1_decode_attn_groups": [vp, i64, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, i64, f32, i32, vp],
    "t1_decode_rope_append": [vp, i64, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, i32, vp],
    "t1_decode_attn": [vp, i64, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, i64, f32, vp],
    "t1_sample": [vp, i64, i32, i32, f32, i32, u64, vp, vp, i32, vp, i32, i32, vp, vp, i32, vp, vp, vp],
    "t1_incr": [vp, i32, vp],
    # sampler.cu
    "t1_sample_topk": [vp, i64, i32, i32, f32, i32, u64, vp, vp, i32, vp, i32, i32, vp, vp, i32, vp, vp, vp, i64, vp],
    "t1_sample_topk_workspace_bytes": [i32, i32],
    # optim.cu
    "t1_sumsq_bf16": [vp, i64, vp, vp],
    "t1_adamw": [vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, vp, f32, f32, vp],
    "t1_bf16_to_f32": [vp, vp, i64, vp],
}
_RESTYPES = {"t1_last_error": C.c_char_p, "t1_launch_count": u64, "t1_count_launch": None, "t1_set_pdl": None, "t1_stream_gemm_tune": None, "t1_stream_gemm_dynamic": None, "t1_stream_gemm_max_ctas": None, "t1_gemm_max_ctas": None, "t1_stream_gemm_debug": None,
             "t1_stream_gemm_workspace_bytes": i64, "t1_sample_topk_workspace_bytes": i64,
             "t1_attn_bwd_tc_workspace_bytes": i64}

GEMM_OUT_F32 = 1
GEMM_ACCUMULATE = 2
GEMM_BIAS_F32 = 4
GEMM_A_MN = 8       # A given as row-major [K, M]
GEMM_B_MN = 16      # B given as row-major [K, N]


class Timer1Error(RuntimeError):
    pass


def register(name: str, argtypes: list, restype=C.c_int) -> None:
    _SIGNATURES[name] = argtypes
    if restype is not C.c_int:
        _RESTYPES[name] = restype


def load(build_if_missing: bool = True):
    """Load libtimer1.so (building it in-tree with nvcc if it is absent). Raises on failure — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise Timer1Error(f"{LIB_PATH} is missing; run `python -m time_r1_b200.build`")
        from . import build as _build

        _build.build()
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.t1_last_error.argtypes = []
    lib.t1_last_error.restype = C.c_char_p
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def exported_symbols() -> list[str]:
    return sorted(_SIGNATURES) + ["t1_last_error"]


def last_error() -> str:
    return (load().t1_last_error() or b"").decode()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise Timer1Error(f"{what} failed (rc={rc}): {last_error()}")


class CallProfiler:
    """Optional CUDA-event timing of every C-ABI call by entry-point name (tools / bench --kernel-times)."""

    def __init__(self):
        self.enabled = False
        self.records: list = []

    def start(self):
        self.enabled, self.records = True, []

    def stop(self):
        self.enabled = False

    def summary(self) -> dict:
        import torch

        torch.cuda.synchronize()
        out: dict = {}
        for name, e0, e1 in self.records:
            d = out.setdefault(name, [0, 0.0])
            d[0] += 1
            d[1] += e0.elapsed_time(e1)
        return dict(sorted(out.items(), key=lambda kv: -kv[1][1]))


CALL_PROFILER = CallProfiler()


def call(name: str, *args) -> None:
    fn = getattr(load(), name)
    if CALL_PROFILER.enabled:
        import torch

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(fn(*args), name)
        e1.record()
        CALL_PROFILER.records.append((name, e0, e1))
        return
    check(fn(*args), name)


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()
