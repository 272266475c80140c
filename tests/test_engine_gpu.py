"""End-to-end parity of the CUDA engine against the CPU fp32 oracle on a toy Qwen2.5-VL / Qwen2-VL config.

Same seeded bf16-representable weights on both sides; tolerances reflect bf16 activations vs an fp32 oracle.
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu



def build(family="qwen2_5_vl", beta=0.04, use_grpo=False, G=4, C=12, fix_vit=True):
    from oracle import qwen_vl_oracle as O
    from time_r1_b200 import config, engine, weights

    cfg = config.toy(family)
    ocfg = O.default_cfg(**{k: v for k, v in cfg.to_dict().items() if k in O.default_cfg()})
    w_or = O.seeded_weights(ocfg, seed=7, std=0.05)
    mw = weights.ModelWeights(cfg, "cuda", train_vit=not fix_vit)
    mw.load_state_dict(w_or)
    run = config.RunConfig(num_generations=G, max_completion_length=C, beta=beta, use_grpo=use_grpo, learning_rate=1e-3, fix_vit=fix_vit)
    eng = engine.Engine(cfg, mw, run)
    rng = np.random.default_rng(3)
    grid = [(2, 8, 12)]
    N = 2 * 8 * 12
    pix = torch.from_numpy(rng.standard_normal((N, cfg.patch_dim), dtype=np.float32))
    pix = pix.to(torch.bfloat16).float()   # bf16-representable pixels: the product casts them to bf16 (HF :1149)
    nv = N // 4
    ids = np.array([5, 6, cfg.vision_start_token_id] + [cfg.video_token_id] * nv + [9] + list(rng.integers(10, 2000, 11)) + [7, 8])
    prompt = engine.Prompt(input_ids=ids, pixel_values=pix.cuda(), grid_thw=grid, second_per_grid_ts=[1.0])
    return O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid


def relerr(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


@pytest.mark.parametrize("family", ["qwen2_5_vl", "qwen2_vl"])
def test_vit_matches_oracle(family):
    O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(family, beta=0.0)
    emb, _, _ = eng.vit_forward(prompt)
    ref = O.vit_forward(w_or, ocfg, pix, grid)
    assert relerr(emb, ref) < 3e-2      # bf16 activations through 3 blocks + merger vs fp32


@pytest.mark.parametrize("family", ["qwen2_5_vl", "qwen2_vl"])
def test_scoring_logp_entropy_match_oracle(family):
    O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(family, beta=0.0)
    G, C = 4, 12
    rng = np.random.default_rng(5)
    comp = torch.from_numpy(rng.integers(10, 2000, (G, C))).int().cuda()
    emb, x_vit, plan = eng.vit_forward(prompt)
    lp, ent, _ = eng.score(prompt, comp, emb)
    lp_ref, ent_ref = O.score_completions(w_or, ocfg, list(ids), comp.cpu().numpy(), pix, grid)
    assert (lp.cpu() - lp_ref).abs().max().item() < 5e-2     # log-prob units; bf16 hidden states, fp32 softmax
    assert (ent.cpu() - ent_ref).abs().max().item() < 5e-2


def test_greedy_rollout_follows_oracle_argmax():
    O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(beta=0.0)
    G, C = 4, 10
    emb, _, _ = eng.vit_forward(prompt)
    comp = eng.rollout(prompt, emb, G, C, temperature=0.0, top_k=0, seed=1)
    comp2 = eng.rollout(prompt, emb, G, C, temperature=0.0, top_k=0, seed=1, use_graph=False)
    assert torch.equal(comp, comp2), "graph replay and eager decode must agree"
    assert (comp == comp[0:1]).all(), "greedy: all G sequences identical"
    # teacher-force the produced tokens through the oracle: each must be (near-)argmax of the oracle logits
    seq = list(ids) + [int(t) for t in comp[0].cpu()]
    logits = O.full_forward_logits(w_or, ocfg, seq, pix, grid)
    P = len(ids)
    for c in range(C):
        row = logits[P - 1 + c]
        tok = int(comp[0, c])
        if tok in cfg.stop_token_ids:
            break
        assert row[tok] >= row.max() - 0.05 * (row.max() - row.min()), f"token {c} is not a near-argmax"


def test_decode_attention_forms_roll_out_the_same_tokens():
    """RunConfig.decode_attn: "cluster", "wide" and "auto" (cluster below decode_attn_wide_min_prompt prompt tokens, wide from
    there on) are the same arithmetic up to the merge order of the split-KV partials: greedy rollouts agree, and "auto" takes
    the form its threshold says."""
    import dataclasses

    from time_r1_b200 import engine

    O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(beta=0.0)
    G, C = 4, 10
    emb, _, _ = eng.vit_forward(prompt)
    outs, forms = {}, {}
    for name, over in (("cluster", dict(decode_attn="cluster")), ("wide", dict(decode_attn="wide")),
                       ("auto_short", dict(decode_attn="auto", decode_attn_wide_min_prompt=10 ** 6)),
                       ("auto_long", dict(decode_attn="auto", decode_attn_wide_min_prompt=16))):
        e2 = engine.Engine(cfg, eng.w, dataclasses.replace(eng.run, **over), with_ref=False)
        outs[name] = e2.rollout(prompt, emb, G, C, temperature=0.0, top_k=0, seed=1).cpu()
        forms[name] = e2._last_decode_state["attn_form"]
    assert forms == dict(cluster="cluster", wide="wide", auto_short="cluster", auto_long="wide")
    for name in ("wide", "auto_short", "auto_long"):
        assert torch.equal(outs[name], outs["cluster"]), name


def test_eager_multi_step_decode_equals_the_captured_graph():
    """use_graph=False runs the decode steps back to back on one stream with programmatic dependent launch: kernels down the chain
    read the previous step's state (step counter, K/V rows) BEFORE their griddepcontrol.wait, so each step's first kernel must be a
    plain stream-order boundary (a graph launch is one by construction).  Regression: sampled eager != sampled graph on the same
    engine before that boundary existed."""
    O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(beta=0.0)
    G, C = 4, 24
    emb, _, _ = eng.vit_forward(prompt)
    a = eng.rollout(prompt, emb, G, C, temperature=1.0, top_k=50, seed=11)
    for _ in range(3):
        b = eng.rollout(prompt, emb, G, C, temperature=1.0, top_k=50, seed=11, use_graph=False)
        assert torch.equal(a, b)
    c = eng.rollout(prompt, emb, G, C, temperature=0.0, top_k=0, seed=1)
    d = eng.rollout(prompt, emb, G, C, temperature=0.0, top_k=0, seed=1, use_graph=False)
    assert torch.equal(c, d)


def test_sampled_rollout_and_decode_logits_consistency():
    """Sampled tokens scored by the packed scoring pass must have the log-probs the decode path sampled from:
    decode (prefix-shared KV + skinny GEMMs) and scoring (packed tcgen05 path) are two routes to the same logits."""
    O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(beta=0.0)
    G, C = 4, 12
    emb, _, _ = eng.vit_forward(prompt)
    comp = eng.rollout(prompt, emb, G, C, temperature=1.0, top_k=50, seed=11)
    assert comp.shape == (G, C) and (comp >= 0).all() and (comp < cfg.vocab_size).all()
    assert not (comp == comp[0:1]).all(), "independent samples should differ"
    lp, _, _ = eng.score(prompt, comp, emb)
    lp_ref, _ = O.score_completions(w_or, ocfg, list(ids), comp.cpu().numpy(), pix, grid)
    assert (lp.cpu() - lp_ref).abs().max().item() < 5e-2
    # every sampled token must be inside the oracle's top-50 (+ slack for bf16 ties at the boundary)
    P = len(ids)
    for g in range(G):
        seq = list(ids) + [int(t) for t in comp[g].cpu()]
        logits = O.full_forward_logits(w_or, ocfg, seq, pix, grid)
        for c in range(C):
            row = logits[P - 1 + c]
            kth = torch.topk(row, 60).values[-1]
            assert row[int(comp[g, c])] >= kth


@pytest.mark.parametrize("family,use_grpo,beta", [("qwen2_5_vl", False, 0.04), ("qwen2_5_vl", True, 0.04), ("qwen2_5_vl", False, 0.0),
                                                  ("qwen2_vl", False, 0.04), ("qwen2_vl", True, 0.0)])
def test_grpo_step_loss_and_gradients_match_oracle(family, use_grpo, beta):
    """Both model families (BASELINE.json names Qwen2-VL, the reference instantiates Qwen2.5-VL): the Qwen2-VL step runs the
    LayerNorm merger backward (t1_layernorm_bwd; HF modeling_qwen2_vl.py:313-326).  Anchored bars for the same inputs live in
    tests/test_parity_7b_width_gpu.py; the fixed bars below are the older, looser ones."""
    O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(family, beta=beta, use_grpo=use_grpo)
    G, C = 4, 12
    rng = np.random.default_rng(9)
    comp = torch.from_numpy(rng.integers(10, 2000, (G, C))).int()
    comp[1, 5] = cfg.eos_token_id
    comp[1, 6:] = cfg.pad_token_id
    rewards = rng.random((G, 2)).astype(np.float32)
    # perturb the policy away from the reference so the KL term is non-trivial
    with torch.no_grad():
        eng.w.train["model.layers.1.mlp.down_proj.weight"].mul_(1.5)
        eng.w.train["visual.merger.mlp.2.weight"].mul_(1.2)
    out = eng.grpo_step(prompt, lambda c: rewards, completion_override=comp)
    # ---- oracle: same computation with autograd on fp32 copies of the (perturbed) policy weights ----
    w_pol = {k: v.clone() for k, v in w_or.items()}
    w_pol["model.layers.1.mlp.down_proj.weight"] = (w_pol["model.layers.1.mlp.down_proj.weight"].to(torch.bfloat16) * 1.5).float()
    w_pol["visual.merger.mlp.2.weight"] = (w_pol["visual.merger.mlp.2.weight"].to(torch.bfloat16) * 1.2).float()
    watch = ["model.layers.0.self_attn.q_proj.weight", "model.layers.0.self_attn.k_proj.bias", "model.layers.1.mlp.gate_proj.weight",
             "model.layers.1.mlp.down_proj.weight", "model.layers.0.input_layernorm.weight", "model.norm.weight",
             "lm_head.weight", "model.embed_tokens.weight", "visual.merger.mlp.0.weight", "visual.merger.mlp.2.bias",
             "visual.merger.ln_q.weight", "model.layers.1.self_attn.o_proj.weight", "model.layers.1.self_attn.v_proj.weight"]
    if family == "qwen2_vl":
        watch.append("visual.merger.ln_q.bias")
    for k in watch:
        w_pol[k].requires_grad_(True)
    lp, ent = O.score_completions(w_pol, ocfg, list(ids), comp.numpy(), pix, grid)
    with torch.no_grad():
        lp_ref, _ = O.score_completions(w_or, ocfg, list(ids), comp.numpy(), pix, grid)
    mask = O.completion_mask(comp.long(), cfg.eos_token_id)
    loss, metrics, adv = O.grpo_loss(lp, lp_ref, mask, torch.from_numpy(rewards), G, beta, use_grpo=use_grpo, entropy=ent.detach())
    loss.backward()
    assert abs(out["loss"].item() - loss.item()) < 2e-2 * max(1.0, abs(loss.item()))
    assert torch.allclose(out["advantages"].cpu(), adv, atol=1e-4, rtol=1e-4)
    assert torch.equal(out["completion_mask"].cpu(), mask)
    if beta:
        assert abs(out["stats"][1].item() - metrics["kl"]) < 2e-2 * max(1.0, metrics["kl"])
    assert abs(out["stats"][2].item() - metrics["generation_entropy"]) < 2e-2
    sd = {}
    from time_r1_b200.weights import ModelWeights
    gw = ModelWeights.__new__(ModelWeights)
    gw.cfg, gw.train, gw.frozen = cfg, eng.grads, eng.w.frozen
    sd = gw.export_state_dict()
    for k in watch:
        g_ref = w_pol[k].grad
        g = sd[k].float().cpu().reshape(g_ref.shape)
        denom = g_ref.abs().max().item() + 1e-12
        err = (g - g_ref).abs().max().item() / denom
        cos = torch.nn.functional.cosine_similarity(g.flatten(), g_ref.flatten(), dim=0).item()
        assert err < 0.08 and cos > 0.99, f"{k}: rel err {err:.3f}, cosine {cos:.4f}"


@pytest.mark.parametrize("family,beta", [("qwen2_5_vl", 0.04), ("qwen2_vl", 0.04), ("qwen2_5_vl", 0.0)])
def test_vit_block_gradients_match_oracle_with_fix_vit_false(family, beta):
    """fix_vit=False (the reference default, main.py:46-49; timer1_trainer.py:272-280 leaves model.visual trainable): the ViT
    patch-embed and every block receive gradients -- windowed + full attention backward at head_dim 80, RMSNorm / LayerNorm,
    SwiGLU / QuickGELU MLP, rotary inverse, the padded MLP columns staying exactly zero -- against autograd on the fp32 oracle.
    Two accumulated micro-steps check the accumulate path as well (gradient == 2 x the single-step gradient)."""
    O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(family, beta=beta, fix_vit=False)
    assert eng.w.frozen.numel == 0 and "visual.blocks.0.attn.qkv.weight" in eng.w.train
    G, C = 4, 12
    rng = np.random.default_rng(9)
    comp = torch.from_numpy(rng.integers(10, 2000, (G, C))).int()
    comp[1, 5] = cfg.eos_token_id
    comp[1, 6:] = cfg.pad_token_id
    rewards = rng.random((G, 2)).astype(np.float32)
    with torch.no_grad():    # policy != reference INSIDE the tower: the reference pass must run its own (unperturbed) ViT
        eng.w.train["visual.blocks.1.attn.proj.weight"].mul_(1.25)
        eng.w.train["model.layers.1.mlp.down_proj.weight"].mul_(1.5)
    out = eng.grpo_step(prompt, lambda c: rewards, completion_override=comp)
    w_pol = {k: v.clone() for k, v in w_or.items()}
    w_pol["visual.blocks.1.attn.proj.weight"] = (w_pol["visual.blocks.1.attn.proj.weight"].to(torch.bfloat16) * 1.25).float()
    w_pol["model.layers.1.mlp.down_proj.weight"] = (w_pol["model.layers.1.mlp.down_proj.weight"].to(torch.bfloat16) * 1.5).float()
    mlp = (["mlp.gate_proj.weight", "mlp.up_proj.bias", "mlp.down_proj.weight", "mlp.down_proj.bias"] if family == "qwen2_5_vl"
           else ["mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias", "norm1.bias", "norm2.bias"])
    watch = ["visual.patch_embed.proj.weight", "visual.merger.mlp.0.weight", "model.layers.0.self_attn.q_proj.weight"]
    for i in range(cfg.vit_depth):
        watch += [f"visual.blocks.{i}.{n}" for n in ["norm1.weight", "norm2.weight", "attn.qkv.weight", "attn.qkv.bias",
                                                     "attn.proj.weight", "attn.proj.bias"] + mlp]
    for k in watch:
        w_pol[k].requires_grad_(True)
    lp, ent = O.score_completions(w_pol, ocfg, list(ids), comp.numpy(), pix, grid)
    with torch.no_grad():
        lp_ref, _ = O.score_completions(w_or, ocfg, list(ids), comp.numpy(), pix, grid)
    mask = O.completion_mask(comp.long(), cfg.eos_token_id)
    loss, metrics, adv = O.grpo_loss(lp, lp_ref, mask, torch.from_numpy(rewards), G, beta, entropy=ent.detach())
    loss.backward()
    assert abs(out["loss"].item() - loss.item()) < 2e-2 * max(1.0, abs(loss.item()))
    if beta:
        assert metrics["kl"] > 1e-4 and abs(out["stats"][1].item() - metrics["kl"]) < 2e-2 * max(1.0, metrics["kl"])
        assert (out["ref_logp"].cpu() - lp_ref).abs().max().item() < 5e-2
    from time_r1_b200.weights import ModelWeights
    gw = ModelWeights.__new__(ModelWeights)
    gw.cfg, gw.train, gw.frozen = cfg, eng.grads, eng.w.frozen
    sd = {k: v.float().cpu().clone() for k, v in gw.export_state_dict().items()}
    for k in watch:
        g_ref = w_pol[k].grad
        g = sd[k].reshape(g_ref.shape)
        err = (g - g_ref).abs().max().item() / (g_ref.abs().max().item() + 1e-12)
        cos = torch.nn.functional.cosine_similarity(g.flatten(), g_ref.flatten(), dim=0).item()
        assert err < 0.08 and cos > 0.99, f"{k}: rel err {err:.3f}, cosine {cos:.4f}"
    if family == "qwen2_5_vl":      # zero-padded MLP rows / columns (3420 -> 3456 in the 7B tower) never receive a gradient
        vi, ip = cfg.vit_intermediate, cfg.vit_inter_padded
        assert ip > vi
        gu, dn = eng.grads["visual.blocks.0.mlp.gate_up.weight"], eng.grads["visual.blocks.0.mlp.down_proj.weight"]
        assert gu[vi:ip].abs().max().item() == 0 and gu[ip + vi:].abs().max().item() == 0 and dn[:, vi:].abs().max().item() == 0
        assert eng.grads["visual.blocks.0.mlp.gate_up.bias"][vi:ip].abs().max().item() == 0
    # accumulation: a second identical micro-step doubles every gradient
    g1 = eng.grads.buf.float().clone()
    eng.grpo_step(prompt, lambda c: rewards, completion_override=comp)
    g2 = eng.grads.buf.float()
    assert torch.allclose(g2, 2 * g1, rtol=2e-2, atol=2e-2 * g1.abs().max().item())
    # the optimizer moves the tower, the reference copy keeps the initial one
    before = eng.w.train["visual.blocks.0.attn.qkv.weight"].clone()
    eng.optimizer_step(lr=1e-3)
    torch.cuda.synchronize()
    assert not torch.equal(eng.w.train["visual.blocks.0.attn.qkv.weight"], before)
    if beta:
        assert torch.equal(eng.ref["visual.blocks.0.attn.qkv.weight"], before)


def test_fix_vit_mismatch_between_run_and_weights_raises():
    from time_r1_b200 import config, engine, weights

    cfg = config.toy()
    mw = weights.ModelWeights(cfg, "cuda")
    with pytest.raises(ValueError, match="train_vit"):
        engine.Engine(cfg, mw, config.RunConfig(num_generations=2, max_completion_length=4, fix_vit=False))


def test_optimizer_step_moves_weights_like_torch_adamw():
    O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(beta=0.04)
    G, C = 4, 12
    rng = np.random.default_rng(2)
    comp = torch.from_numpy(rng.integers(10, 2000, (G, C))).int()
    rewards = rng.random((G, 2)).astype(np.float32)
    eng.grpo_step(prompt, lambda c: rewards, completion_override=comp)
    g = eng.grads.buf.float().clone()
    p0 = eng.master.clone()
    norm = eng.optimizer_step(lr=1e-3)
    torch.cuda.synchronize()
    assert abs(norm.item() - g.norm().item()) < 1e-3 * g.norm().item()
    ref = torch.nn.Parameter(p0.clone())
    ref.grad = g * min(1.0, 1.0 / (g.norm().item() + 1e-6))
    torch.optim.AdamW([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0).step()
    assert torch.allclose(eng.master, ref.detach(), atol=1e-6, rtol=1e-4)
    assert torch.equal(eng.w.train.buf, eng.master.to(torch.bfloat16))
    assert eng.grad_fresh


def test_rollout_stops_at_eos_pads_and_masks_like_the_reference():
    """HF generate semantics the trainer relies on (timer1_trainer.py:568-590): a sequence that emits a stop token keeps it,
    every later position is pad_token_id, the completion mask covers tokens up to AND INCLUDING the first EOS, and the
    decode loop ends early once every sequence has finished."""
    import dataclasses

    from time_r1_b200 import config, engine, ops, weights

    O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(beta=0.0)
    G, C = 4, 40
    emb, _, _ = eng.vit_forward(prompt)
    base = eng.rollout(prompt, emb, G, C, temperature=0.0, top_k=0, seed=1)[0].cpu().tolist()
    k = next(i for i in range(1, C) if base[i] not in base[:i])      # first step whose token did not occur earlier
    stop_tok = base[k]
    cfg2 = dataclasses.replace(cfg, eos_token_id=stop_tok, stop_token_ids=(stop_tok, cfg.pad_token_id))
    mw = weights.ModelWeights(cfg2, "cuda")
    mw.load_state_dict(w_or)
    eng2 = engine.Engine(cfg2, mw, config.RunConfig(num_generations=G, max_completion_length=C, beta=0.0), with_ref=False)
    emb2, _, _ = eng2.vit_forward(prompt)
    comp = eng2.rollout(prompt, emb2, G, C, temperature=0.0, top_k=0, seed=1)
    for g in range(G):
        row = comp[g].cpu().tolist()
        assert row[:k + 1] == base[:k + 1]
        assert all(t == cfg2.pad_token_id for t in row[k + 1:]), row
    mask, lengths = ops.eos_mask(comp, cfg2.eos_token_id)
    assert mask.sum(1).tolist() == [k + 1] * G and lengths.tolist() == [k + 1] * G
    # early exit: the step counter stopped at the first all-finished check (check_every = 32), not at C - 1
    assert int(eng2._last_decode_state["step"].item()) <= 32


@pytest.mark.parametrize("sampled", [False, True])
def test_kept_activation_levels_give_the_same_gradients(sampled):
    """Engine._save_level: 0 = recompute everything from the layer-input checkpoints, 1 = keep gate_up, 2 = also keep q/k/v,
    attention output and LSE.  cfg3 / cfg5 fall back to the lower levels when HBM runs short: every level must produce the
    same loss and the same gradients (the kept tensors ARE what the recompute would produce)."""
    from time_r1_b200 import config, engine, weights

    res = []
    for level in (0, 1, 2):
        O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(beta=0.04)
        eng.save_gate_up = level
        G, C = 4, 12
        rng = np.random.default_rng(9)
        rewards = rng.random((G, 2)).astype(np.float32)
        if sampled:      # rollout path: the prefill's kept rows are reused by the policy scoring pass
            out = eng.grpo_step(prompt, lambda c: rewards, seed=5)
        else:
            comp = torch.from_numpy(rng.integers(10, 2000, (G, C))).int()
            out = eng.grpo_step(prompt, lambda c: rewards, completion_override=comp)
        torch.cuda.synchronize()
        res.append((out["loss"].item(), out["completion_ids"].cpu().clone(), out["logp"].cpu().clone(), eng.grads.buf.float().cpu().clone()))
    for level in (1, 2):
        assert torch.equal(res[level][1], res[0][1]), "same seed -> same completions at every level"
        assert abs(res[level][0] - res[0][0]) < 1e-6
        assert torch.allclose(res[level][2], res[0][2], atol=1e-5)
        g0, g = res[0][3], res[level][3]
        cos = torch.nn.functional.cosine_similarity(g0.double(), g.double(), dim=0).item()
        assert cos > 0.99999 and (g - g0).abs().max().item() <= 2.0 ** -7 * g0.abs().max().item(), (level, cos)


def test_fused_head_step_equals_the_unfused_one():
    """RunConfig.fused_head (lm_head + log-softmax statistics / softmax backward in the GEMM epilogue, no logits in HBM) gives the
    same log-probs, entropy, loss and gradients as the GEMM -> fp32 logits -> t1_logprob_entropy / t1_logits_backward chain."""
    res = []
    for fused in (False, True):
        O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(beta=0.04)
        eng.run.fused_head = fused
        G, C = 4, 12
        rng = np.random.default_rng(9)
        comp = torch.from_numpy(rng.integers(10, 2000, (G, C))).int()
        rewards = rng.random((G, 2)).astype(np.float32)
        out = eng.grpo_step(prompt, lambda c: rewards, completion_override=comp)
        torch.cuda.synchronize()
        res.append((out["loss"].item(), out["logp"].cpu().clone(), out["entropy"].cpu().clone(), eng.grads.buf.float().cpu().clone()))
    assert abs(res[0][0] - res[1][0]) < 1e-6
    assert torch.allclose(res[0][1], res[1][1], atol=2e-5) and torch.allclose(res[0][2], res[1][2], atol=1e-4)
    g0, g1 = res[0][3], res[1][3]
    assert torch.nn.functional.cosine_similarity(g0.double(), g1.double(), dim=0).item() > 0.99999
    assert (g1 - g0).abs().max().item() <= 2.0 ** -6 * g0.abs().max().item()
