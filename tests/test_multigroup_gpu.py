"""Joint rollout of several prompt-groups (the micro-batches of one optimizer step) must equal separate rollouts."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_multi_group_rollout_matches_single_group_greedy_and_grads_accumulate():
    from time_r1_b200 import config, engine, weights

    cfg = config.toy()
    mw = weights.ModelWeights(cfg, "cuda")
    mw.init_random(seed=3, std=0.05)
    G, C = 4, 9
    run = config.RunConfig(num_generations=G, max_completion_length=C, beta=0.04)
    eng = engine.Engine(cfg, mw, run)
    rng = np.random.default_rng(0)
    prompts = []
    for grid in [(2, 8, 12), (2, 8, 8)]:
        N = grid[0] * grid[1] * grid[2]
        ids = np.array([5, 6, cfg.vision_start_token_id] + [cfg.video_token_id] * (N // 4) + [9] + list(rng.integers(10, 2000, 5 + len(prompts) * 3)) + [7])
        pix = torch.from_numpy(rng.standard_normal((N, cfg.patch_dim)).astype(np.float32)).cuda()
        prompts.append(engine.Prompt(input_ids=ids, pixel_values=pix, grid_thw=[grid], second_per_grid_ts=[1.0]))
    embs = [eng.vit_forward(p)[0] for p in prompts]
    joint = eng.rollout_multi(prompts, embs, G, C, temperature=0.0, top_k=0, seed=1)
    joint_eager = eng.rollout_multi(prompts, embs, G, C, temperature=0.0, top_k=0, seed=1, use_graph=False)
    for gi, p in enumerate(prompts):
        single = eng.rollout(p, embs[gi], G, C, temperature=0.0, top_k=0, seed=1)
        assert torch.equal(joint[gi], single), f"group {gi}: joint decode differs from single-group decode"
        assert torch.equal(joint_eager[gi], single)
    sampled = eng.rollout_multi(prompts, embs, G, C, temperature=1.0, top_k=50, seed=5)
    assert not torch.equal(sampled[0], sampled[1])
    # gradient accumulation over the two groups == sum of the per-group gradients.  Distinct random completions:
    # with identical (greedy) completions and zero-sum advantages the true gradient cancels to rounding noise.
    rw = [rng.random((G, 2)).astype(np.float32) for _ in prompts]
    joint = [torch.from_numpy(rng.integers(10, 2000, (G, C))).int().cuda() for _ in prompts]
    eng.grpo_multi_step(prompts, [lambda c, r=r: r for r in rw], completion_overrides=joint, loss_scale=0.5)
    g_joint = eng.grads.buf.float().clone()
    eng.grad_fresh = True
    eng.grpo_step(prompts[0], lambda c: rw[0], completion_override=joint[0], loss_scale=0.5)
    g0 = eng.grads.buf.float().clone()
    eng.grad_fresh = True
    eng.grpo_step(prompts[1], lambda c: rw[1], completion_override=joint[1], loss_scale=0.5)
    g1 = eng.grads.buf.float().clone()
    ref = g0 + g1
    err = (g_joint - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)
    assert err < 2e-2, f"accumulated gradient differs from the sum of per-group gradients: {err}"


def test_prefill_reuse_equals_full_policy_scoring():
    """Policy scoring that reuses the rollout's prefill for the prompt rows must reproduce the full packed scoring pass
    (log-probs, loss and gradients)."""
    from time_r1_b200 import config, engine, weights

    cfg = config.toy()
    mw = weights.ModelWeights(cfg, "cuda")
    mw.init_random(seed=4, std=0.05)
    G, C = 4, 9
    rng = np.random.default_rng(1)
    grid = (2, 8, 12)
    N = grid[0] * grid[1] * grid[2]
    ids = np.array([5, 6, cfg.vision_start_token_id] + [cfg.video_token_id] * (N // 4) + [9] + list(rng.integers(10, 2000, 7)) + [7])
    pix = torch.from_numpy(rng.standard_normal((N, cfg.patch_dim)).astype(np.float32)).cuda()
    prompt = engine.Prompt(input_ids=ids, pixel_values=pix, grid_thw=[grid], second_per_grid_ts=[1.0])
    rw = rng.random((G, 2)).astype(np.float32)
    eng = engine.Engine(cfg, mw, config.RunConfig(num_generations=G, max_completion_length=C, beta=0.04, reuse_prefill=True))
    out_r = eng.grpo_step(prompt, lambda c: rw, seed=3)                    # rollout + reuse path
    g_r = eng.grads.buf.float().clone()
    eng.grad_fresh = True
    out_f = eng.grpo_step(prompt, lambda c: rw, completion_override=out_r["completion_ids"])   # full scoring path
    g_f = eng.grads.buf.float().clone()
    assert (out_r["logp"] - out_f["logp"]).abs().max().item() < 2e-2
    assert abs(out_r["loss"].item() - out_f["loss"].item()) < 2e-3
    err = (g_r - g_f).abs().max().item() / (g_f.abs().max().item() + 1e-12)
    cos = torch.nn.functional.cosine_similarity(g_r, g_f, dim=0).item()
    assert err < 3e-2 and cos > 0.999, (err, cos)


def test_vit_prefetch_on_side_stream_changes_nothing():
    """The frozen ViT blocks of the next step's prompts may run on a side stream under the optimizer phase
    (Engine.prefetch_vit): a step fed by the prefetch handle produces bit-identical results."""
    from test_engine_gpu import build

    outs = []
    for use_prefetch in (False, True):
        O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(beta=0.04)
        rewards = np.random.default_rng(0).random((4, 2)).astype(np.float32)
        handle = eng.prefetch_vit([prompt]) if use_prefetch else None
        out = eng.grpo_multi_step([prompt], [lambda c: rewards], seed=3, vit_prefetch=handle)[0]
        gn = eng.optimizer_step()
        torch.cuda.synchronize()
        outs.append((out["completion_ids"].clone(), out["logp"].clone(), out["loss"].clone(), gn.clone()))
    for a, b in zip(outs[0][:3], outs[1][:3]):
        assert torch.equal(a, b)
    # the gradient norm goes through fp32 atomics (norm-weight gradients): equal up to summation order
    assert abs(outs[0][3].item() - outs[1][3].item()) <= 1e-3 * outs[0][3].item()


def test_saved_gate_up_gives_the_same_gradients_as_recompute():
    """RunConfig.save_gate_up keeps the gate_up projection of every layer from the forward passes (prefill rows from the
    rollout, completion rows from the scoring pass) so the backward skips that GEMM: same log-probs, same gradients."""
    import dataclasses

    from test_engine_gpu import build
    from time_r1_b200 import engine

    res = []
    for save in (False, True):
        O, cfg, ocfg, w_or, eng, prompt, pix, ids, grid = build(beta=0.04)
        eng.save_gate_up = save
        rewards = np.random.default_rng(0).random((4, 2)).astype(np.float32)
        out = eng.grpo_multi_step([prompt], [lambda c: rewards], seed=3)[0]
        torch.cuda.synchronize()
        res.append((out["completion_ids"].clone(), out["logp"].clone(), eng.grads.buf.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    a, b = res[0][2].float(), res[1][2].float()
    # bf16 gradient buffer: the two routes may differ by an ulp on a few elements (fp32 atomics in the norm backward feed dx)
    assert (a - b).abs().max().item() <= 1e-2 * a.abs().max().item(), ((a - b).abs().max().item(), a.abs().max().item())
    assert torch.nn.functional.cosine_similarity(a, b, dim=0).item() > 0.9999
