"""Model / run configuration for the Qwen2-VL and Qwen2.5-VL families (one struct, two ViT variants).

Field meanings follow HF `Qwen2_5_VLConfig` / `Qwen2VLConfig` (configuration_qwen2_5_vl.py); the trainer in the
reference hard-codes Qwen2.5-VL (src/time_r1/rl/timer1_trainer.py:244,298) while BASELINE.json names Qwen2-VL, so
both are first-class here.
"""
from __future__ import annotations

import os
from dataclasses import asdict, dataclass, field, replace


@dataclass
class ModelConfig:
    family: str = "qwen2_5_vl"  # or "qwen2_vl"
    # ---- LLM ----
    hidden_size: int = 3584
    intermediate_size: int = 18944
    num_hidden_layers: int = 28
    num_attention_heads: int = 28
    num_key_value_heads: int = 4
    head_dim: int = 128
    vocab_size: int = 152064
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1e6
    mrope_section: tuple = (16, 24, 24)
    tie_word_embeddings: bool = False
    # ---- ViT ----
    vit_hidden: int = 1280
    vit_depth: int = 32
    vit_heads: int = 16
    vit_intermediate: int = 3420
    patch_size: int = 14
    temporal_patch_size: int = 2
    spatial_merge_size: int = 2
    window_size: int = 112
    fullatt_block_indexes: tuple = (7, 15, 23, 31)
    in_channels: int = 3
    tokens_per_second: int = 2
    # ---- special tokens ----
    video_token_id: int = 151656
    image_token_id: int = 151655
    vision_start_token_id: int = 151652
    eos_token_id: int = 151645
    pad_token_id: int = 151643
    stop_token_ids: tuple = (151645, 151643)

    @property
    def is_q25(self) -> bool:
        return self.family == "qwen2_5_vl"

    @property
    def vit_head_dim(self) -> int:
        return self.vit_hidden // self.vit_heads

    @property
    def vit_inter_padded(self) -> int:
        """ViT MLP width padded to a multiple of 64 (zero rows/cols: exact) so TMA row strides stay 16-byte aligned."""
        return (self.vit_intermediate + 63) // 64 * 64

    @property
    def patch_dim(self) -> int:
        return self.in_channels * self.temporal_patch_size * self.patch_size ** 2

    @property
    def merge_unit(self) -> int:
        return self.spatial_merge_size ** 2

    @property
    def qkv_dim(self) -> int:
        return (self.num_attention_heads + 2 * self.num_key_value_heads) * self.head_dim

    def to_dict(self) -> dict:
        return asdict(self)


def qwen2_5_vl_7b() -> ModelConfig:
    return ModelConfig()


def qwen2_5_vl_3b() -> ModelConfig:
    return ModelConfig(hidden_size=2048, intermediate_size=11008, num_hidden_layers=36, num_attention_heads=16,
                       num_key_value_heads=2, vocab_size=151936, tie_word_embeddings=True)


def qwen2_vl_7b() -> ModelConfig:
    return ModelConfig(family="qwen2_vl", vit_intermediate=5120, fullatt_block_indexes=(), window_size=0,
                       tokens_per_second=1)


def qwen2_vl_2b() -> ModelConfig:
    return ModelConfig(family="qwen2_vl", hidden_size=1536, intermediate_size=8960, num_hidden_layers=28,
                       num_attention_heads=12, num_key_value_heads=2, vocab_size=151936, tie_word_embeddings=True,
                       vit_intermediate=5120, fullatt_block_indexes=(), window_size=0, tokens_per_second=1)


def toy(family: str = "qwen2_5_vl", **over) -> ModelConfig:
    """Small config the CPU oracle finishes in seconds (used by the parity tests and smoke())."""
    base = dict(family=family, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                num_key_value_heads=1, head_dim=128, vocab_size=2048, vit_hidden=160, vit_depth=3, vit_heads=2,
                vit_intermediate=200, fullatt_block_indexes=(1,), video_token_id=2040, image_token_id=2041,
                vision_start_token_id=2042, eos_token_id=2043, pad_token_id=2044, stop_token_ids=(2043, 2044))
    if family == "qwen2_vl":
        base.update(fullatt_block_indexes=(), window_size=0, tokens_per_second=1, vit_intermediate=320)
    base.update(over)
    return ModelConfig(**base)


PRESETS = {
    "qwen2.5-vl-7b": qwen2_5_vl_7b, "qwen2.5-vl-3b": qwen2_5_vl_3b, "qwen2-vl-7b": qwen2_vl_7b,
    "qwen2-vl-2b": qwen2_vl_2b, "toy": toy,
}


def from_hf_config(d: dict) -> ModelConfig:
    """Build from an HF config.json dict (Qwen2VLConfig / Qwen2_5_VLConfig, flat or with text_config)."""
    t = d.get("text_config", d)
    v = d.get("vision_config", {})
    fam = "qwen2_5_vl" if "2_5" in d.get("model_type", "qwen2_5_vl") else "qwen2_vl"
    rope = t.get("rope_scaling") or t.get("rope_parameters") or {}
    cfg = ModelConfig(
        family=fam, hidden_size=t["hidden_size"], intermediate_size=t["intermediate_size"],
        num_hidden_layers=t["num_hidden_layers"], num_attention_heads=t["num_attention_heads"],
        num_key_value_heads=t["num_key_value_heads"], head_dim=t["hidden_size"] // t["num_attention_heads"],
        vocab_size=t["vocab_size"], rms_norm_eps=t.get("rms_norm_eps", 1e-6), rope_theta=t.get("rope_theta", 1e6),
        mrope_section=tuple(rope.get("mrope_section", (16, 24, 24))),
        tie_word_embeddings=d.get("tie_word_embeddings", t.get("tie_word_embeddings", False)),
        # Qwen2-VL: embed_dim is the ViT width and vision hidden_size the merger's output width; Qwen2.5-VL: hidden_size is the width
        vit_hidden=v.get("embed_dim", 1280) if fam == "qwen2_vl" else v.get("hidden_size", 1280), vit_depth=v.get("depth", 32),
        vit_heads=v.get("num_heads", 16),
        vit_intermediate=v.get("intermediate_size", int(v.get("embed_dim", 1280) * v.get("mlp_ratio", 4))),
        patch_size=v.get("patch_size", 14), temporal_patch_size=v.get("temporal_patch_size", 2),
        spatial_merge_size=v.get("spatial_merge_size", 2), window_size=v.get("window_size", 0 if fam == "qwen2_vl" else 112),
        fullatt_block_indexes=tuple(v.get("fullatt_block_indexes", ())), tokens_per_second=v.get("tokens_per_second", 2 if fam == "qwen2_5_vl" else 1),
        video_token_id=d.get("video_token_id", 151656), image_token_id=d.get("image_token_id", 151655),
        vision_start_token_id=d.get("vision_start_token_id", 151652),
    )
    return cfg


@dataclass
class RunConfig:
    """Knobs of one GRPO step (defaults = scripts/posttrain/train_rl.sh + trl 0.17 GRPOConfig defaults)."""
    num_generations: int = 8
    max_completion_length: int = 200
    temperature: float = 1.0
    top_k: int = 50
    beta: float = 0.04
    epsilon_low: float = 0.2   # timer1_trainer.py:388-393 hard-codes 0.2 / 0.2
    epsilon_high: float = 0.2
    use_grpo: bool = False
    learning_rate: float = 1e-6
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_eps: float = 1e-8
    weight_decay: float = 0.0
    max_grad_norm: float = 1.0
    gradient_accumulation_steps: int = 1
    fix_vit: bool = True
    ref_quirk_mrope: bool = True   # scoring pass omits second_per_grid_ts (timer1_trainer.py:601-607)
    seed: int = 42
    logit_chunk_rows: int = 2048
    # lm_head fused with log-softmax / gather / entropy (forward) and with the softmax backward (recompute) in the tcgen05 GEMM
    # epilogue: fp32 logits never reach HBM (t1_gemm_lse_parts + t1_lse_merge, t1_gemm_dlogits); False = GEMM -> fp32 logits chunk
    # -> t1_logprob_entropy / t1_logits_backward (the A/B arm)
    fused_head: bool = os.environ.get("T1_FUSED_HEAD", "1") != "0"
    # what the forward passes keep for the backward besides the layer inputs: 0 / False = nothing (recompute all),
    # 1 = gate_up projections, 2 / True = + q/k/v, attention output, LSE; "auto" = as much as the free HBM allows
    save_gate_up: object = "auto"
    reuse_prefill: bool = True     # policy scoring reuses the rollout's prefill (same weights) for the prompt rows
    # decode step: "stream" = tcgen05 stream-K weight streaming + grouped attention launches chained with programmatic
    # dependent launch; "legacy" = HMMA skinny GEMM + per-group attention on forked graph branches
    decode_impl: str = "stream"
    decode_pdl: bool = True
    # decode attention inside the "stream" step: "cluster" = one launch, split-KV partials merged through distributed
    # shared memory (64 CTAs at cfg2); "wide" = the same CTA program without clusters on every SM (144 CTAs at cfg2), partials
    # merged through an L2-resident workspace behind a device-side arrival counter; "split" = split-KV kernel + combine
    # kernel through a global workspace (HMMA, round 1)
    # kernel (HMMA, round 1); "auto" (default) = cluster below decode_attn_wide_min_prompt prompt tokens, wide from there on:
    # measured per decode step (2 groups, 7B): cfg2 (P 4216) cluster 3.260 / wide 3.279 ms, cfg3 (P 8312) 3.394 / 3.383,
    # cfg5 (P 16504) 3.665 / 3.546 (profiles/r02_decode_wide_vs_cluster_by_context.txt)
    decode_attn: str = os.environ.get("T1_DECODE_ATTN", "auto")
    # at the round's final kernels (slab-aligned o_proj on 112 CTAs, hoisted loads) the wide form wins at cfg2's context as well:
    # P 4216 cluster 3.139 / wide 3.115 ms, P 8312 3.279 / 3.208, P 16504 3.545 / 3.369 (profiles/r02_decode_step_final_by_context.txt,
    # r02_decode_wide_vs_cluster_cfg2_final.txt); shorter prompts than measured stay on clusters
    decode_attn_wide_min_prompt: int = 4096
    # fuse the residual add + RMSNorm that follows the o / down projection into that stream GEMM (5 kernels per layer
    # instead of 7).  Measured neutral on B200 (3.53 vs 3.49 ms per step): what looked like kernel-boundary latency is
    # the finishing-time spread of the 148 streaming CTAs, which the fused tail has to wait out as well.
    decode_fuse_norm: bool = False
    # fold the residual add + RMSNorm between two decode GEMMs into their epilogues (t1_stream_gemm_fold / _scaled): the
    # producer's finished slabs update the residual stream and emit z = norm_w * resid plus per-slab row sums of squares, the
    # consumer streams against z and multiplies its accumulator by rsqrt(mean + eps).  No norm kernel, no grid-wide wait:
    # 147 instead of 202 kernels per decode step, 3.35 -> 3.23 ms per step on B200 (unlike the fused tail above, nothing
    # waits out the CTA finishing spread twice).  Same math as HF up to where the bf16 roundings fall.
    decode_fold_norm: bool = True
    # with the folded norms: launch o_proj + gate_up and down_proj + the next layer's qkv as ONE kernel each
    # (t1_stream_gemm_fold_pair: device-side arrival counter between the two GEMMs, the weight stream never stops at it);
    # 92 instead of 147 kernels per decode step, bit-identical tokens.  Measured (profiles/r02_decode_pair_ab.txt,
    # r02_decode_timeline_pair.txt): the counter is passed 7.8 / 6.2 us after CTA 0's phase A -- no earlier than the kernel
    # boundary it replaces (7.3 / 7.0 us: both wait for the slowest CTA's fix-up, not for the launch machinery) -- and the
    # two-phase producer issues its loads slower (gate_up 49 vs 40 us): 3.66 vs 3.20 ms per decode step.  OFF by default.
    decode_pair: bool = os.environ.get("T1_DECODE_PAIR", "0") == "1"
    # CTAs of the decode step's o_proj GEMM: 0 = one per SM, -1 = only as many as the cluster-form attention launch leaves free
    # (148 - 8 * Hkv * groups), n > 0 = that many
    decode_o_ctas: int = int(os.environ.get("T1_DECODE_O_CTAS", "0"))
    decode_qkv_ctas: int = int(os.environ.get("T1_DECODE_QKV_CTAS", "0"))
    decode_down_ctas: int = int(os.environ.get("T1_DECODE_DOWN_CTAS", "0"))
    # MB of each layer's gate_up weights pulled into L2 (t1_stream_gemm_prefetch_next) by the idle warps of the qkv ("qkv")
    # or o_proj ("o") GEMM of the same layer while the chain is latency-bound (qkv -> attention -> o); 0 = off
    decode_prefetch_mb: int = int(os.environ.get("T1_DECODE_PF_MB", "0"))
    decode_prefetch_at: str = os.environ.get("T1_DECODE_PF_AT", "o")
    # overlap the prompt rows of the reference-model scoring pass (tensor-bound, independent of the sampled tokens) with the
    # decode (HBM-bound): the decode's weight-streaming GEMMs keep to `decode_ctas` SMs on a high-priority stream, the
    # reference prefill runs on a side stream with its GEMMs capped at `side_ctas` CTAs.  0 = off.
    side_ctas: int = 0       # experimental: 32 to enable
    decode_ctas: int = 116
