"""ORACLE-side CPU baseline (bench.py `cpu_baseline` leg and `--impl reference` arm only).

Times the reference's own arithmetic for one GRPO step on the host cores with the fp32 oracle port
(qwen_vl_oracle.py): *reference-style* — for each of the G copies the ViT and the prompt are recomputed in generate,
in policy scoring and in reference scoring, logits are produced at every position (src/time_r1/rl/
timer1_trainer.py:452-460,568-573,594-632), then autograd backward and one torch.optim.AdamW step (lr 1e-6,
betas (0.9, 0.999), eps 1e-8, wd 0, clip 1.0 — HF Trainer defaults the reference runs with).

The full cfg2 step (7B, 4.2k-token prompt) is hours of CPU time, so each timed step is a BOUNDED SAMPLE of that
workload: the 7B layer/ViT-block/merger dimensions with depth 1+1, a (2,16,16) patch grid (128 vision tokens), G=2,
C=4 and a 32768-row vocabulary; its measured time is scaled to the full step by the ratio of reference-style
FLOPs (time_r1_b200.flops.reference_style_flops for the full config / the sample config).  The sample description is
returned with the number.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from . import qwen_vl_oracle as O


def _randn_weights(cfg, seed=0, std=0.02):
    """Same tensor set as qwen_vl_oracle.seeded_weights but generated with torch (fast for 7B-sized matrices)."""
    g = torch.Generator().manual_seed(seed)
    q25 = cfg["family"] == "qwen2_5_vl"
    w = {}
    for n, shp in _shapes(cfg).items():
        if q25 and (".mlp.fc" in n or n.endswith("norm1.bias") or n.endswith("norm2.bias") or n.endswith("ln_q.bias")):
            continue
        if not q25 and n.startswith("visual.blocks") and ("gate_proj" in n or "up_proj" in n or "down_proj" in n):
            continue
        if cfg["tie_word_embeddings"] and n == "lm_head.weight":
            continue
        is_norm_w = n.endswith(("norm1.weight", "norm2.weight", "layernorm.weight", "model.norm.weight", "ln_q.weight"))
        w[n] = torch.ones(shp) if is_norm_w else torch.randn(shp, generator=g) * std
    return w


def _shapes(cfg):
    D, I = cfg["hidden_size"], cfg["intermediate_size"]
    Hq, Hkv, hd, V = cfg["num_attention_heads"], cfg["num_key_value_heads"], cfg["head_dim"], cfg["vocab_size"]
    vd, vi = cfg["vit_hidden"], cfg["vit_intermediate"]
    kdim = cfg["in_channels"] * cfg["temporal_patch_size"] * cfg["patch_size"] ** 2
    m = cfg["spatial_merge_size"] ** 2
    s = {"visual.patch_embed.proj.weight": (vd, kdim), "visual.merger.ln_q.weight": (vd,), "visual.merger.ln_q.bias": (vd,),
         "visual.merger.mlp.0.weight": (vd * m, vd * m), "visual.merger.mlp.0.bias": (vd * m,),
         "visual.merger.mlp.2.weight": (D, vd * m), "visual.merger.mlp.2.bias": (D,),
         "model.embed_tokens.weight": (V, D), "model.norm.weight": (D,), "lm_head.weight": (V, D)}
    for i in range(cfg["vit_depth"]):
        p = f"visual.blocks.{i}."
        s.update({p + "norm1.weight": (vd,), p + "norm1.bias": (vd,), p + "norm2.weight": (vd,), p + "norm2.bias": (vd,),
                  p + "attn.qkv.weight": (3 * vd, vd), p + "attn.qkv.bias": (3 * vd,), p + "attn.proj.weight": (vd, vd),
                  p + "attn.proj.bias": (vd,), p + "mlp.gate_proj.weight": (vi, vd), p + "mlp.gate_proj.bias": (vi,),
                  p + "mlp.up_proj.weight": (vi, vd), p + "mlp.up_proj.bias": (vi,), p + "mlp.down_proj.weight": (vd, vi),
                  p + "mlp.down_proj.bias": (vd,), p + "mlp.fc1.weight": (vi, vd), p + "mlp.fc1.bias": (vi,),
                  p + "mlp.fc2.weight": (vd, vi), p + "mlp.fc2.bias": (vd,)})
    for i in range(cfg["num_hidden_layers"]):
        p = f"model.layers.{i}."
        s.update({p + "input_layernorm.weight": (D,), p + "post_attention_layernorm.weight": (D,),
                  p + "self_attn.q_proj.weight": (Hq * hd, D), p + "self_attn.q_proj.bias": (Hq * hd,),
                  p + "self_attn.k_proj.weight": (Hkv * hd, D), p + "self_attn.k_proj.bias": (Hkv * hd,),
                  p + "self_attn.v_proj.weight": (Hkv * hd, D), p + "self_attn.v_proj.bias": (Hkv * hd,),
                  p + "self_attn.o_proj.weight": (D, Hq * hd), p + "mlp.gate_proj.weight": (I, D),
                  p + "mlp.up_proj.weight": (I, D), p + "mlp.down_proj.weight": (D, I)})
    return s


class ReferenceStyleStep:
    """One reference-style GRPO step on the CPU (bounded sample)."""

    def __init__(self, cfg: dict, grid, G: int, C: int, n_text: int = 16, seed: int = 42, threads: int | None = None):
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.cfg, self.grid, self.G, self.C = cfg, grid, G, C
        self.w = _randn_weights(cfg, seed)
        self.w_ref = {k: v.clone() for k, v in self.w.items() if not k.startswith("visual.blocks") and not k.startswith("visual.patch")}
        for k, v in self.w.items():
            if k.startswith("visual.blocks") or k.startswith("visual.patch"):
                self.w_ref[k] = v                      # frozen ViT shared (fix_vit)
            else:
                v.requires_grad_(True)
        self.trainable = [v for v in self.w.values() if v.requires_grad]
        self.opt = torch.optim.AdamW(self.trainable, lr=1e-6, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
        rng = np.random.default_rng(seed)
        N = sum(t * h * w for t, h, w in grid)
        kdim = cfg["in_channels"] * cfg["temporal_patch_size"] * cfg["patch_size"] ** 2
        self.pix = torch.from_numpy(rng.standard_normal((N, kdim), dtype=np.float32))
        nv = N // cfg["spatial_merge_size"] ** 2
        V = cfg["vocab_size"]
        self.ids = [1, 2, 3, cfg["vision_start_token_id"]] + [cfg["video_token_id"]] * nv + [4] + list(rng.integers(10, V - 100, n_text)) + [5, 6]
        self.P = len(self.ids)
        self.rng = rng

    def step(self) -> float:
        cfg, G, C = self.cfg, self.G, self.C
        t0 = time.perf_counter()
        with torch.no_grad():   # HF generate: G expanded copies, each with its own ViT + prefill + decode
            comp = O.generate(self.w, cfg, self.ids, self.pix, self.grid, G, C, temperature=1.0, top_k=50,
                              seed=int(self.rng.integers(1 << 30)), share_prefix=False)
        mask = O.completion_mask(comp, cfg["eos_token_id"])
        lps, ents = [], []
        for g in range(G):      # policy scoring: ViT on repeated pixels + logits at every position, with grad
            lp, ent = O.score_completions(self.w, cfg, self.ids, comp[g:g + 1].numpy(), self.pix, self.grid)
            lps.append(lp[0]); ents.append(ent[0])
        lp = torch.stack(lps)
        with torch.no_grad():
            rl = torch.stack([O.score_completions(self.w_ref, cfg, self.ids, comp[g:g + 1].numpy(), self.pix, self.grid)[0][0] for g in range(G)])
        rewards = torch.from_numpy(self.rng.random((G, 2)).astype(np.float32))
        loss, _, _ = O.grpo_loss(lp, rl, mask, rewards, G, beta=0.04)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.trainable, 1.0)
        self.opt.step()
        return time.perf_counter() - t0


SAMPLE_DESC = ("oracle port, reference-style arithmetic (ViT+prompt recomputed per copy, logits at all positions), "
               "7B layer dims with 1 ViT block + 1 decoder layer, grid (2,16,16), G=2, C=4, vocab 32768, fp32, "
               "scaled to the full cfg2 step by reference-style FLOPs")


def sample_cfg(full_cfg: dict) -> dict:
    c = dict(full_cfg)
    c.update(num_hidden_layers=1, vit_depth=1, fullatt_block_indexes=(0,), vocab_size=32768,
             video_token_id=32760, image_token_id=32761, vision_start_token_id=32762, eos_token_id=32763, pad_token_id=32764)
    return c
