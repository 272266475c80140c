"""ORACLE tooling — pins oracle/qwen_vl_oracle.py against the HF classes installed in THIS container
(transformers 5.5.0; the reference pins 4.51.1 and executes the same modules through timer1_trainer.py:244,452-457).
Writes tests/golden/hf_toy_<family>.npz: vision-tower output and LM logits of a seeded toy model.  Position ids are fed
explicitly (the oracle's 4.51.1-semantics M-RoPE indices) because 5.5.0's own index builder changed.

    python oracle/gen_golden_model.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import qwen_vl_oracle as O  # noqa: E402
from time_r1_b200 import config  # noqa: E402


def hf_model(family, cfg):
    if family == "qwen2_5_vl":
        from transformers import Qwen2_5_VLConfig as C, Qwen2_5_VLForConditionalGeneration as M
        vis = dict(depth=cfg.vit_depth, hidden_size=cfg.vit_hidden, intermediate_size=cfg.vit_intermediate, num_heads=cfg.vit_heads,
                   patch_size=cfg.patch_size, temporal_patch_size=cfg.temporal_patch_size, spatial_merge_size=cfg.spatial_merge_size,
                   window_size=cfg.window_size, fullatt_block_indexes=list(cfg.fullatt_block_indexes), out_hidden_size=cfg.hidden_size,
                   tokens_per_second=cfg.tokens_per_second, hidden_act="silu")
    else:
        from transformers import Qwen2VLConfig as C, Qwen2VLForConditionalGeneration as M
        vis = dict(depth=cfg.vit_depth, embed_dim=cfg.vit_hidden, hidden_size=cfg.hidden_size, mlp_ratio=cfg.vit_intermediate // cfg.vit_hidden,
                   num_heads=cfg.vit_heads, patch_size=cfg.patch_size, temporal_patch_size=cfg.temporal_patch_size,
                   spatial_merge_size=cfg.spatial_merge_size, hidden_act="quick_gelu")
    text = dict(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size, num_hidden_layers=cfg.num_hidden_layers,
                num_attention_heads=cfg.num_attention_heads, num_key_value_heads=cfg.num_key_value_heads, vocab_size=cfg.vocab_size,
                rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta, max_position_embeddings=32768,
                rope_scaling={"type": "default", "mrope_section": list(cfg.mrope_section), "rope_type": "default"},
                tie_word_embeddings=cfg.tie_word_embeddings, use_sliding_window=False)
    hc = C(text_config=text, vision_config=vis, video_token_id=cfg.video_token_id, image_token_id=cfg.image_token_id,
           vision_start_token_id=cfg.vision_start_token_id, tie_word_embeddings=cfg.tie_word_embeddings)
    hc._attn_implementation = "eager"
    m = M(hc).float().eval()
    return m


def load_weights(m, w, cfg):
    sd = m.state_dict()
    new = {}
    for k, v in w.items():
        hk = k
        if k.startswith("visual."):
            hk = "model." + k
        elif k.startswith("model."):
            hk = "model.language_model." + k[len("model."):]
        if hk == "model.visual.patch_embed.proj.weight":
            v = v.reshape(sd[hk].shape)
        assert hk in sd, hk
        assert sd[hk].shape == v.shape, (hk, sd[hk].shape, v.shape)
        new[hk] = v
    missing = set(sd) - set(new)
    if cfg.tie_word_embeddings:
        missing -= {"lm_head.weight"}
    assert not missing, sorted(missing)[:5]
    m.load_state_dict(new, strict=not cfg.tie_word_embeddings)


def main():
    for family in ("qwen2_5_vl", "qwen2_vl"):
        cfg = config.toy(family)
        ocfg = O.default_cfg(**{k: v for k, v in cfg.to_dict().items() if k in O.default_cfg()})
        w = O.seeded_weights(ocfg, seed=7, std=0.05)
        m = hf_model(family, cfg)
        load_weights(m, w, cfg)
        rng = np.random.default_rng(3)
        grid = [(2, 8, 12)]
        N = 2 * 8 * 12
        pix = torch.from_numpy(rng.standard_normal((N, cfg.patch_dim), dtype=np.float32))
        ids = [5, 6, cfg.vision_start_token_id] + [cfg.video_token_id] * (N // 4) + [9, 10, 11, 12, 7, 8] + [33, 44, 55]
        g = torch.tensor(grid)
        with torch.no_grad():
            vout = m.model.visual(pix, grid_thw=g)
            emb = vout.pooler_output if hasattr(vout, "pooler_output") else vout
            pos, delta = O.mrope_positions(ocfg, ids, grid, [1.0])
            out = m(input_ids=torch.tensor([ids]), pixel_values_videos=pix, video_grid_thw=g, position_ids=pos[:, None, :],
                    attention_mask=torch.ones(1, len(ids), dtype=torch.long))
        logits = out.logits[0]
        # oracle on the same inputs
        emb_o = O.vit_forward(w, ocfg, pix, grid)
        logits_o = O.full_forward_logits(w, ocfg, ids, pix, grid, [1.0])
        print(family, "vit max|d|", (emb - emb_o).abs().max().item(), "logits max|d|", (logits - logits_o).abs().max().item(),
              "scale", logits.abs().max().item())
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"hf_toy_{family}.npz"), vision_embeds=emb.numpy(),
                            logits_tail=logits[-8:].numpy(), ids=np.array(ids), grid=np.array(grid), pos=pos.numpy(),
                            transformers_version=np.array(__import__("transformers").__version__))


if __name__ == "__main__":
    main()
