"""Algorithmic work model of one GRPO step: FLOPs = 2 x MACs of the dense contractions plus
4 * heads*head_dim * sum(kv_len) per attention layer.  Used by bench.py for roofline fractions and for scaling the
bounded CPU sample to the full workload."""
from __future__ import annotations

from .config import ModelConfig


def linear_params(cfg: ModelConfig) -> dict:
    D, I, L = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers
    qd, kd = cfg.num_attention_heads * cfg.head_dim, cfg.num_key_value_heads * cfg.head_dim
    vd, vi = cfg.vit_hidden, cfg.vit_intermediate
    m = cfg.merge_unit
    vit_block = 3 * vd * vd + vd * vd + (3 * vd * vi if cfg.is_q25 else 2 * vd * vi)
    return dict(
        vit_per_patch=cfg.vit_depth * vit_block + vd * cfg.patch_dim,
        merger_per_token=(vd * m) * (vd * m) + cfg.hidden_size * vd * m,
        llm_per_token=L * (D * (qd + 2 * kd) + qd * D + 3 * D * I),
        head_per_token=cfg.vocab_size * D,
    )


def vit_attn_flops(cfg: ModelConfig, grid_thw) -> float:
    """4 * d_model * sum over query rows of visible keys, per block (windowed vs full)."""
    vd = cfg.vit_hidden
    total = 0.0
    for t, h, w in grid_thw:
        n_slot = h * w
        full = 4.0 * vd * t * n_slot * n_slot
        win = 4.0 * vd * t * n_slot * 64.0   # 8x8-patch windows (upper bound for ragged edges)
        n_full = len(cfg.fullatt_block_indexes) if cfg.is_q25 else cfg.vit_depth
        total += n_full * full + (cfg.vit_depth - n_full) * win
    return total


def step_flops(cfg: ModelConfig, grid_thw, P: int, G: int, C: int, with_ref: bool = True, recompute: bool = True) -> dict:
    """Exact-math minimum (prefix sharing, ViT once, logits at completion rows only)."""
    lp = linear_params(cfg)
    N = sum(t * h * w for t, h, w in grid_thw)
    Nv = N // cfg.merge_unit
    L, qd = cfg.num_hidden_layers, cfg.num_attention_heads * cfg.head_dim
    vit = 2.0 * N * lp["vit_per_patch"] + 2.0 * Nv * lp["merger_per_token"] + vit_attn_flops(cfg, grid_thw)
    attn_prefill = 4.0 * qd * L * (P * (P + 1) / 2)
    prefill = 2.0 * P * lp["llm_per_token"] + attn_prefill + 2.0 * lp["head_per_token"]
    dec_tokens = G * (C - 1)
    attn_dec = 4.0 * qd * L * G * sum(P + c + 1 for c in range(C - 1))
    decode = 2.0 * dec_tokens * (lp["llm_per_token"] + lp["head_per_token"]) + attn_dec
    T = P + G * C
    attn_score = attn_prefill + 4.0 * qd * L * G * sum(P + c + 1 for c in range(C))
    score = 2.0 * T * lp["llm_per_token"] + attn_score + 2.0 * G * C * lp["head_per_token"]
    bwd = 2.0 * score + (score if recompute else 0.0)
    total = vit + prefill + decode + score + bwd + (score + 2.0 * Nv * lp["merger_per_token"] if with_ref else 0.0)
    return dict(vit=vit, prefill=prefill, decode=decode, score_fwd=score, backward=bwd, total=total)


def reference_style_flops(cfg: ModelConfig, grid_thw, P: int, G: int, C: int, with_ref: bool = True,
                          recompute: bool = True) -> float:
    """What the reference's arithmetic costs for the same step: ViT and prompt repeated for each of the G copies in
    generate, policy scoring and reference scoring; logits at every position (timer1_trainer.py:452-460,594-596)."""
    lp = linear_params(cfg)
    N = sum(t * h * w for t, h, w in grid_thw)
    Nv = N // cfg.merge_unit
    L, qd = cfg.num_hidden_layers, cfg.num_attention_heads * cfg.head_dim
    vit = 2.0 * N * lp["vit_per_patch"] + 2.0 * Nv * lp["merger_per_token"] + vit_attn_flops(cfg, grid_thw)
    attn_seq = 4.0 * qd * L * ((P + C) * (P + C + 1) / 2)
    gen = G * (vit + 2.0 * P * lp["llm_per_token"] + 4.0 * qd * L * (P * (P + 1) / 2) + 2.0 * lp["head_per_token"]
               + 2.0 * (C - 1) * (lp["llm_per_token"] + lp["head_per_token"])
               + 4.0 * qd * L * sum(P + c + 1 for c in range(C - 1)))
    score = G * (vit + 2.0 * (P + C) * (lp["llm_per_token"] + lp["head_per_token"]) + attn_seq)
    bwd = 2.0 * score + (score if recompute else 0.0)
    return gen + score + bwd + (score if with_ref else 0.0)
