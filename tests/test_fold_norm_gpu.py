"""RMSNorm folded into the decode GEMMs around it (t1_stream_gemm_fold / t1_stream_gemm_scaled): kernel-level checks
against the unfolded kernels and torch fp32, a captured PDL chain, and the decode step of the engine in fold mode against
the CPU oracle."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16 = torch.bfloat16


def rnd(*shape, seed=0, scale=1.0, dtype=BF16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


def close(a, b, atol, rtol, what=""):
    a, b = a.float(), b.float()
    err = (a - b).abs()
    bad = (err > atol + rtol * b.abs()).sum().item()
    assert bad == 0, f"{what}: {bad} elements out of tolerance, max err {err.max().item():.4g}"


@pytest.mark.parametrize("M,N,K,N2", [(16, 3584, 3584, 4608), (16, 3584, 18944, 1024), (5, 256, 512, 384), (40, 1024, 264, 640), (64, 128, 1024, 200)])
def test_fold_producer_and_scaled_consumer(M, N, K, N2):
    from time_r1_b200 import ops

    eps = 1e-6
    MP = 16 if M <= 16 else (32 if M <= 32 else 64)
    x, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05)
    resid0, nw = rnd(M, N, seed=3), rnd(N, seed=4).abs() + 0.5
    w2, b2 = rnd(N2, N, seed=5, scale=0.05), rnd(N2, seed=6)
    ws = ops.stream_gemm_workspace(M, DEV)
    # ---- producer: residual stream, z and the per-slab row sums ----
    y_ref = ops.stream_gemm(x, w, ws)                                    # same kernel, same stream-K cut points
    r_ref = (resid0.float() + y_ref.float()).to(BF16)
    z_ref = (nw.float() * r_ref.float()).to(BF16)
    for _ in range(2):                                                   # twice: counters reset themselves
        resid, z = resid0.clone(), torch.full((M, N), float("nan"), device=DEV, dtype=BF16)
        parts = torch.full((N // 128, MP), float("nan"), device=DEV)
        ops.stream_gemm_fold(x, w, ws, resid, nw, z, parts)
        assert torch.equal(resid, r_ref), "residual stream must equal bf16(resid + bf16(x W^T))"
        assert torch.equal(z, z_ref), "z must equal bf16(norm_w * resid)"
        want_parts = r_ref.float().pow(2).view(M, N // 128, 128).sum(-1).t()
        close(parts[:, :M], want_parts, 1e-4, 1e-5, "per-slab row sums of squares")
    assert int(ws[:65536].view(torch.int32).abs().sum().item()) == 0, "slab counters must be left at zero"
    # ---- consumer: row factor applied in the epilogue ----
    rfac = torch.rsqrt(r_ref.float().pow(2).mean(-1, keepdim=True) + eps)
    want = rfac * (z_ref.float() @ w2.float().t()) + b2.float()
    got = ops.stream_gemm_scaled(z, w2, ws, parts, N // 128, eps, bias=b2)
    again = ops.stream_gemm_scaled(z, w2, ws, parts, N // 128, eps, bias=b2)
    assert torch.equal(got, again)
    close(got, want, 2e-2, 1e-2, "scaled consumer vs fp32")
    # and it is the unfolded computation up to bf16 rounding order: W2 . rmsnorm(resid) with the norm kernel in between
    h = ops.rmsnorm(r_ref.clone(), nw, eps)
    unfolded = ops.stream_gemm(h, w2, ws, bias=b2)
    scale_ref = want.abs().max().item()
    assert (got.float() - unfolded.float()).abs().max().item() < 2e-2 * scale_ref + 2e-2
    # ---- SwiGLU consumer ----
    if N2 % 128 == 0:
        gotsw = ops.stream_gemm_scaled(z, w2, ws, parts, N // 128, eps, swiglu=True)
        lin = rfac * (z_ref.float() @ w2.float().t())
        close(gotsw, torch.nn.functional.silu(lin[:, : N2 // 2]) * lin[:, N2 // 2:], 3e-2, 3e-2, "scaled swiglu consumer vs fp32")
    assert int(ws[:65536].view(torch.int32).abs().sum().item()) == 0


@pytest.mark.parametrize("M,N,K,N2,swiglu", [(16, 3584, 3584, 2 * 18944, True), (16, 3584, 18944, 4608, False), (5, 256, 512, 384, False),
                                             (40, 1024, 264, 640, True), (64, 128, 1024, 200, False), (16, 512, 128, 2 * 1024, True),
                                             (8, 3584, 3584, 256, False)])
def test_fold_pair_is_bit_identical_to_the_two_launches(M, N, K, N2, swiglu):
    """t1_stream_gemm_fold_pair (both GEMMs in one launch, device-side arrival counter between them) == t1_stream_gemm_fold
    followed by t1_stream_gemm_scaled, bit for bit: residual stream, z, row sums and y; repeated launches find the counters
    re-armed; the separate kernels still work on the same workspace afterwards."""
    from time_r1_b200 import ops

    eps = 1e-6
    MP = 16 if M <= 16 else (32 if M <= 32 else 64)
    x, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05)
    resid0, nw = rnd(M, N, seed=3), rnd(N, seed=4).abs() + 0.5
    w2 = rnd(N2, N, seed=5, scale=0.05)
    b2 = None if swiglu else rnd(N2, seed=6)
    ws = ops.stream_gemm_workspace(M, DEV)

    def separate():
        resid, z, parts = resid0.clone(), torch.zeros(M, N, device=DEV, dtype=BF16), torch.zeros(N // 128, MP, device=DEV)
        ops.stream_gemm_fold(x, w, ws, resid, nw, z, parts)
        y = ops.stream_gemm_scaled(z, w2, ws, parts, N // 128, eps, bias=b2, swiglu=swiglu)
        return resid, z, parts, y

    want = separate()
    for _ in range(3):
        resid, z = resid0.clone(), torch.full((M, N), float("nan"), device=DEV, dtype=BF16)
        parts = torch.full((N // 128, MP), float("nan"), device=DEV)
        y = ops.stream_gemm_fold_pair(x, w, ws, resid, nw, z, parts, w2, eps, bias=b2, swiglu=swiglu)
        torch.cuda.synchronize()
        assert torch.equal(resid, want[0]) and torch.equal(z, want[1])
        assert torch.equal(parts[:, :M], want[2][:, :M])
        assert torch.equal(y, want[3]), (y.float() - want[3].float()).abs().max().item()
        assert int(ws[:65536].view(torch.int32).abs().sum().item()) == 0, "slab and pair counters must be left at zero"
    again = separate()
    assert all(torch.equal(a[:, :M] if a.dim() == 2 and a.shape[1] == MP else a, b[:, :M] if b.dim() == 2 and b.shape[1] == MP else b)
               for a, b in zip(again, want))


def test_paired_chain_under_graph_with_pdl():
    """[pair(o + gate_up) -> pair(down + qkv)] x 4 captured with programmatic dependent launch == the four-launch folded chain
    run eagerly (the pair kernels re-arm their counters behind griddepcontrol.wait of the next launch)."""
    from time_r1_b200 import lib, ops

    M, D, I = 16, 512, 1024
    x0, a0 = rnd(M, D, seed=1), rnd(M, D, seed=7)
    wn1, wn2 = rnd(D, seed=2).abs() + 0.5, rnd(D, seed=8).abs() + 0.5
    wo, w1, w2, wq = rnd(D, D, seed=9, scale=0.05), rnd(2 * I, D, seed=3, scale=0.05), rnd(D, I, seed=4, scale=0.05), rnd(D, D, seed=5, scale=0.05)
    bq = rnd(D, seed=11)
    ws = ops.stream_gemm_workspace(M, DEV)
    z, act, attn = torch.empty_like(x0), torch.empty(M, I, device=DEV, dtype=BF16), torch.empty_like(x0)
    parts = torch.zeros(2, D // 128, 64, device=DEV)
    x = x0.clone()

    def chain(paired):
        for _ in range(4):
            if paired:
                ops.stream_gemm_fold_pair(attn, wo, ws, x, wn1, z, parts[0], w1, 1e-6, out=act, swiglu=True)
                ops.stream_gemm_fold_pair(act, w2, ws, x, wn2, z, parts[1], wq, 1e-6, out=attn, bias=bq)
            else:
                ops.stream_gemm_fold(attn, wo, ws, x, wn1, z, parts[0])
                ops.stream_gemm_scaled(z, w1, ws, parts[0], D // 128, 1e-6, out=act, swiglu=True)
                ops.stream_gemm_fold(act, w2, ws, x, wn2, z, parts[1])
                ops.stream_gemm_scaled(z, wq, ws, parts[1], D // 128, 1e-6, out=attn, bias=bq)

    attn.copy_(a0)
    chain(False)
    torch.cuda.synchronize()
    want_x, want_attn = x.clone(), attn.clone()
    x.copy_(x0); attn.copy_(a0)
    lib.load().t1_set_pdl(1)
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            chain(True)                              # warm-up on the capture stream
            x.copy_(x0); attn.copy_(a0)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                chain(True)
        torch.cuda.current_stream().wait_stream(s)
    finally:
        lib.load().t1_set_pdl(0)
    for _ in range(2):
        x.copy_(x0); attn.copy_(a0)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(x, want_x) and torch.equal(attn, want_attn)
    assert torch.isfinite(x.float()).all()


def test_folded_chain_under_graph_with_pdl():
    """[fold -> scaled swiglu -> fold -> scaled] x 4 captured with programmatic dependent launch == the same chain run
    eagerly without it (z / parts / the residual stream are read only after the dependency resolves)."""
    from time_r1_b200 import lib, ops

    M, D, I = 16, 512, 1024
    x0, a0 = rnd(M, D, seed=1), rnd(M, D, seed=7)
    wn1, wn2 = rnd(D, seed=2).abs() + 0.5, rnd(D, seed=8).abs() + 0.5
    wo, w1, w2, wq = rnd(D, D, seed=9, scale=0.05), rnd(2 * I, D, seed=3, scale=0.05), rnd(D, I, seed=4, scale=0.05), rnd(D, D, seed=5, scale=0.05)
    ws = ops.stream_gemm_workspace(M, DEV)
    z, act, attn = torch.empty_like(x0), torch.empty(M, I, device=DEV, dtype=BF16), torch.empty_like(x0)
    parts = torch.zeros(2, D // 128, 64, device=DEV)
    x = x0.clone()

    def chain():
        for _ in range(4):
            ops.stream_gemm_fold(attn, wo, ws, x, wn1, z, parts[0])
            ops.stream_gemm_scaled(z, w1, ws, parts[0], D // 128, 1e-6, out=act, swiglu=True)
            ops.stream_gemm_fold(act, w2, ws, x, wn2, z, parts[1])
            ops.stream_gemm_scaled(z, wq, ws, parts[1], D // 128, 1e-6, out=attn)

    attn.copy_(a0)
    chain()
    torch.cuda.synchronize()
    want_x, want_attn = x.clone(), attn.clone()
    x.copy_(x0); attn.copy_(a0)
    lib.load().t1_set_pdl(1)
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            chain()                                  # warm-up on the capture stream
            x.copy_(x0); attn.copy_(a0)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                chain()
        torch.cuda.current_stream().wait_stream(s)
    finally:
        lib.load().t1_set_pdl(0)
    x.copy_(x0); attn.copy_(a0)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(x, want_x) and torch.equal(attn, want_attn)
    assert torch.isfinite(x.float()).all()


def test_engine_decode_in_fold_mode_follows_the_oracle():
    """Greedy rollout with the folded norms: every produced token is a (near-)argmax of the fp32 oracle's logits, graph
    replay equals eager, and the sampled rollout's tokens score like the oracle says (same bars as the unfolded path)."""
    from test_engine_gpu import build

    O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(beta=0.0)
    eng.run.decode_fold_norm = True
    G, C = 4, 10
    emb, _, _ = eng.vit_forward(prompt)
    comp = eng.rollout(prompt, emb, G, C, temperature=0.0, top_k=0, seed=1)
    comp2 = eng.rollout(prompt, emb, G, C, temperature=0.0, top_k=0, seed=1, use_graph=False)
    assert torch.equal(comp, comp2), "graph replay and eager decode must agree"
    assert (comp == comp[0:1]).all(), "greedy: all G sequences identical"
    seq = list(ids) + [int(t) for t in comp[0].cpu()]
    logits = O.full_forward_logits(w_or, ocfg, seq, pix, grid)
    P = len(ids)
    for c in range(C):
        row = logits[P - 1 + c]
        tok = int(comp[0, c])
        if tok in cfg.stop_token_ids:
            break
        assert row[tok] >= row.max() - 0.05 * (row.max() - row.min()), f"token {c} is not a near-argmax"
    comp = eng.rollout(prompt, emb, G, 12, temperature=1.0, top_k=50, seed=11)
    for g in range(G):
        seq = list(ids) + [int(t) for t in comp[g].cpu()]
        logits = O.full_forward_logits(w_or, ocfg, seq, pix, grid)
        for c in range(12):
            row = logits[P - 1 + c]
            assert row[int(comp[g, c])] >= torch.topk(row, 60).values[-1]


def test_engine_decode_paired_launches_give_the_same_tokens():
    """RunConfig.decode_pair: o_proj + gate_up and down_proj + next qkv as one launch each.  Bit-identical logits, hence the
    same SAMPLED tokens as the five-launches-per-layer chain, with 2 * L - 1 fewer kernels per decode step."""
    import dataclasses

    from test_engine_gpu import build
    from time_r1_b200 import engine

    O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(beta=0.0)
    G, C = 4, 24
    emb, _, _ = eng.vit_forward(prompt)
    res = {}
    for pair in (False, True):
        e2 = engine.Engine(cfg, eng.w, dataclasses.replace(eng.run, decode_fold_norm=True, decode_pair=pair), with_ref=False)
        comp = e2.rollout(prompt, emb, G, C, temperature=1.0, top_k=50, seed=11)
        launches = e2._last_decode_state["launches_per_step"]
        greedy = e2.rollout(prompt, emb, G, C, temperature=0.0, top_k=0, seed=1)
        e3 = engine.Engine(cfg, eng.w, dataclasses.replace(eng.run, decode_fold_norm=True, decode_pair=pair), with_ref=False)
        greedy_eager = e3.rollout(prompt, emb, G, C, temperature=0.0, top_k=0, seed=1, use_graph=False)
        assert torch.equal(greedy, greedy_eager), "graph replay and eager decode must agree"
        res[pair] = (comp.cpu(), launches, greedy.cpu())
    assert torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][2], res[False][2])
    assert res[False][1] - res[True][1] == 2 * cfg.num_hidden_layers - 1
