"""tcgen05 GEMM vs torch fp32 reference on the same bf16 inputs (tolerance: bf16 output rounding)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [
    (128, 256, 64), (128, 256, 256), (256, 512, 128), (100, 256, 64), (300, 264, 200), (1600, 1280, 1176),
    (1000, 3584, 3584), (5836, 4608, 3584), (77, 1536, 8960), (4096, 3584, 5120), (520, 37888 // 8, 3584),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_matches_fp32(M, N, K):
    from time_r1_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    ref = a.float() @ b.float().t()
    out = ops.gemm(a, b)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 1e-2 * scale + 1e-3, f"max err {err} vs scale {scale}"
    out32 = ops.gemm(a, b, out_dtype=torch.float32)
    err32 = (out32 - ref).abs().max().item()
    assert err32 <= 2e-5 * scale * (K ** 0.5) + 1e-4, f"fp32-out max err {err32} vs scale {scale}"


def test_gemm_bias_accumulate_strided():
    from time_r1_b200 import ops

    M, N, K = 333, 512, 320
    g = torch.Generator(device="cuda").manual_seed(5)
    abuf = torch.randn(M, K + 64, device="cuda", generator=g).to(torch.bfloat16)
    a = abuf[:, 8:8 + K]  # strided view, 16-byte aligned start
    b = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g).to(torch.bfloat16)
    ref = a.float() @ b.float().t() + bias.float()
    out = ops.gemm(a, b, bias=bias, out_dtype=torch.float32)
    assert (out - ref).abs().max().item() < 1e-3
    out2 = out.clone()
    ops.gemm(a, b, out=out2, accumulate=True)
    ref2 = ref + a.float() @ b.float().t()
    assert (out2 - ref2).abs().max().item() < 2e-3
    obf = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, out=obf, bias=bias.float())
    assert (obf.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()


@pytest.mark.parametrize("a_t,b_t", [(False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 128), (3584, 4608, 5888), (1000, 3584, 3584), (200, 264, 1000),
                                   (152064 // 8, 3584, 1664), (5888, 3584, 18944 // 4)])
def test_gemm_mn_major_operands(M, N, K, a_t, b_t):
    """Backward-pass forms: A and/or B handed over as row-major [K, rows] and consumed as MN-major UMMA operands
    (dgrad: dX = dY @ W with B = W in place; wgrad: dW = dY^T @ X with both in place) vs fp32 on the same inputs."""
    from time_r1_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M + N * 3 + K * 5)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    ref = a.float() @ b.float().t()
    a_in = a.t().contiguous() if a_t else a
    b_in = b.t().contiguous() if b_t else b
    out32 = ops.gemm(a_in, b_in, out_dtype=torch.float32, a_t=a_t, b_t=b_t)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err32 = (out32 - ref).abs().max().item()
    assert err32 <= 2e-5 * scale * (K ** 0.5) + 1e-4, f"fp32-out max err {err32} vs scale {scale}"
    # same bits as the K-major path on explicitly transposed copies (identical products, identical k order)
    assert torch.equal(out32, ops.gemm(a, b, out_dtype=torch.float32))
    acc = out32.clone()
    ops.gemm(a_in, b_in, out=acc, accumulate=True, a_t=a_t, b_t=b_t)
    assert (acc - 2 * ref).abs().max().item() <= 4e-5 * scale * (K ** 0.5) + 2e-4


@pytest.mark.parametrize("M,I,K,with_bias", [(5888, 18944, 3584, False), (300, 512, 264, True), (16384 // 4, 3456, 1280, True), (77, 200, 72, False)])
def test_gemm_swiglu_epilogue(M, I, K, with_bias):
    """gate_up projection with the SwiGLU in the GEMM epilogue == gemm (bf16) followed by swiglu, bit for bit; the optional
    copy of the projection equals the plain GEMM output."""
    from time_r1_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M + I + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(2 * I, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    b = torch.randn(2 * I, device="cuda", generator=g).to(torch.bfloat16) if with_bias else None
    gu_ref = ops.gemm(a, w, bias=b)
    want = ops.swiglu(gu_ref)
    gu = torch.full_like(gu_ref, float("nan"))
    got = ops.gemm_swiglu(a, w, bias=b, gate_up_out=gu)
    assert torch.equal(got, want), (got.float() - want.float()).abs().max().item()
    assert torch.equal(gu, gu_ref)
    assert torch.equal(ops.gemm_swiglu(a, w, bias=b), want)


@pytest.mark.parametrize("M,N,K", [(130, 72, 64), (257, 200, 128), (1000, 1280, 1176), (96, 3584, 256)])
def test_gemm_tma_store_epilogue_strided_output_and_neighbours(M, N, K):
    """bf16 outputs leave through TMA tensor stores ([32 x 64] boxes): rows >= M and columns >= N are clipped by the TMA unit —
    the columns NEXT to a strided output view (the other projections of a fused buffer) and the rows after it stay untouched."""
    from time_r1_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g).to(torch.bfloat16)
    buf = torch.full((M + 40, N + 136), 7.0, device="cuda", dtype=torch.bfloat16)
    out = buf[8:8 + M, 64:64 + N]
    ops.gemm(a, b, out=out, bias=bias)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t() + bias.float()
    assert (out.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item() + 1e-3
    guard = buf.clone()
    guard[8:8 + M, 64:64 + N] = 7.0
    assert (guard == 7.0).all(), "the TMA store wrote outside the output view"
