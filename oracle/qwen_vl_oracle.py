"""ORACLE (test infrastructure only) — CPU fp32 restatement of the Qwen2-VL / Qwen2.5-VL forward, HF-4.51.1-style
M-RoPE indexing, reference-style sampling and the `compute_loss` numerics of Time-R1.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product (time_r1_b200/) never does.

The arithmetic of the hot path lives in third-party code that is NOT under /root/reference:
transformers==4.51.1 (docs/INSTALL.md:11; call sites src/time_r1/rl/timer1_trainer.py:244,452-457,569-573).  This
file restates its published algorithm following the copy installed here (transformers 5.5.0,
models/qwen2_5_vl/modeling_qwen2_5_vl.py = "Q25", models/qwen2_vl/modeling_qwen2_vl.py = "Q2"), pinning the 4.51.1
M-RoPE index semantics (mirrors vllm/model_executor/models/qwen2_5_vl.py get_mrope_input_positions).
Parity pin: tests/golden/hf_toy_*.npz are outputs of the installed HF classes on a seeded toy config
(oracle/gen_golden_model.py); tests/test_oracle_cpu.py checks this restatement against them.  No reference test pins
model outputs ("parity unpinned" upstream) — the HF goldens are the strongest available anchor.

Weights use the released-checkpoint names ("visual.*", "model.*", "lm_head.weight").
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# configuration helpers
# ------------------------------------------------------------------------------------------------
def default_cfg(**over):
    cfg = dict(
        family="qwen2_5_vl",
        # LLM
        hidden_size=3584, intermediate_size=18944, num_hidden_layers=28, num_attention_heads=28,
        num_key_value_heads=4, head_dim=128, vocab_size=152064, rms_norm_eps=1e-6, rope_theta=1e6,
        mrope_section=(16, 24, 24), tie_word_embeddings=False,
        # ViT
        vit_hidden=1280, vit_depth=32, vit_heads=16, vit_intermediate=3420, patch_size=14, temporal_patch_size=2,
        spatial_merge_size=2, window_size=112, fullatt_block_indexes=(7, 15, 23, 31), in_channels=3,
        tokens_per_second=2,
        # tokens
        video_token_id=151656, image_token_id=151655, vision_start_token_id=151652, eos_token_id=151645,
        pad_token_id=151643,
    )
    cfg.update(over)
    return cfg


def seeded_weights(cfg, seed=0, std=0.02):
    """Deterministic random-init weights from numpy's PCG64 (portable across machines), rounded to bf16 so that
    the bf16 product and the fp32 oracle see identical parameters. Norm weights are 1 + small noise."""
    rng = np.random.default_rng(seed)
    w = {}

    def mat(name, *shape, scale=std):
        a = rng.standard_normal(shape, dtype=np.float32) * scale
        w[name] = torch.from_numpy(a).to(torch.bfloat16).float()

    def norm(name, n, bias=False):
        w[name + ".weight"] = torch.from_numpy(1.0 + 0.1 * rng.standard_normal(n, dtype=np.float32)).to(torch.bfloat16).float()
        if bias:
            mat(name + ".bias", n)

    D, I, L = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"]
    Hq, Hkv, hd, V = cfg["num_attention_heads"], cfg["num_key_value_heads"], cfg["head_dim"], cfg["vocab_size"]
    vd, vi, vl = cfg["vit_hidden"], cfg["vit_intermediate"], cfg["vit_depth"]
    q25 = cfg["family"] == "qwen2_5_vl"
    kdim = cfg["in_channels"] * cfg["temporal_patch_size"] * cfg["patch_size"] ** 2
    mat("visual.patch_embed.proj.weight", vd, kdim)
    for i in range(vl):
        p = f"visual.blocks.{i}."
        norm(p + "norm1", vd, bias=not q25)
        norm(p + "norm2", vd, bias=not q25)
        mat(p + "attn.qkv.weight", 3 * vd, vd)
        mat(p + "attn.qkv.bias", 3 * vd)
        mat(p + "attn.proj.weight", vd, vd)
        mat(p + "attn.proj.bias", vd)
        if q25:
            for nm in ("gate_proj", "up_proj"):
                mat(p + f"mlp.{nm}.weight", vi, vd)
                mat(p + f"mlp.{nm}.bias", vi)
            mat(p + "mlp.down_proj.weight", vd, vi)
            mat(p + "mlp.down_proj.bias", vd)
        else:
            mat(p + "mlp.fc1.weight", vi, vd)
            mat(p + "mlp.fc1.bias", vi)
            mat(p + "mlp.fc2.weight", vd, vi)
            mat(p + "mlp.fc2.bias", vd)
    m = cfg["spatial_merge_size"] ** 2
    norm("visual.merger.ln_q", vd, bias=not q25)
    mat("visual.merger.mlp.0.weight", vd * m, vd * m)
    mat("visual.merger.mlp.0.bias", vd * m)
    mat("visual.merger.mlp.2.weight", D, vd * m)
    mat("visual.merger.mlp.2.bias", D)
    mat("model.embed_tokens.weight", V, D)
    for i in range(L):
        p = f"model.layers.{i}."
        norm(p + "input_layernorm", D)
        norm(p + "post_attention_layernorm", D)
        mat(p + "self_attn.q_proj.weight", Hq * hd, D)
        mat(p + "self_attn.q_proj.bias", Hq * hd)
        mat(p + "self_attn.k_proj.weight", Hkv * hd, D)
        mat(p + "self_attn.k_proj.bias", Hkv * hd)
        mat(p + "self_attn.v_proj.weight", Hkv * hd, D)
        mat(p + "self_attn.v_proj.bias", Hkv * hd)
        mat(p + "self_attn.o_proj.weight", D, Hq * hd)
        mat(p + "mlp.gate_proj.weight", I, D)
        mat(p + "mlp.up_proj.weight", I, D)
        mat(p + "mlp.down_proj.weight", D, I)
    norm("model.norm", D)
    if not cfg["tie_word_embeddings"]:
        mat("lm_head.weight", V, D)
    return w


# ------------------------------------------------------------------------------------------------
# ViT  (Q25:455-518; Q2:771-790)
# ------------------------------------------------------------------------------------------------
def rms_norm(x, w, eps):
    # Q25:66-71 (fp32 here; the bf16 down-cast before the weight multiply is a precision detail)
    return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def vit_rot_pos(cfg, grid_thw):
    """(hpos, wpos) per patch in merge-block order (Q25:382-409)."""
    ms = cfg["spatial_merge_size"]
    out = []
    for t, h, w in grid_thw:
        hp = torch.arange(h).unsqueeze(1).expand(-1, w).reshape(h // ms, ms, w // ms, ms).permute(0, 2, 1, 3).flatten()
        wp = torch.arange(w).unsqueeze(0).expand(h, -1).reshape(h // ms, ms, w // ms, ms).permute(0, 2, 1, 3).flatten()
        out.append(torch.stack([hp, wp], -1).repeat(t, 1))
    return torch.cat(out, 0)


def vit_window_index(cfg, grid_thw):
    """window_index over merged tokens + cumulative window lengths in patches (Q25:411-451, :476)."""
    ms, ps = cfg["spatial_merge_size"], cfg["patch_size"]
    vws = cfg["window_size"] // ms // ps
    unit = ms * ms
    window_index, cu, base = [], [0], 0
    for t, h, w in grid_thw:
        lh, lw = h // ms, w // ms
        index = torch.arange(t * lh * lw).reshape(t, lh, lw)
        pad_h, pad_w = vws - lh % vws, vws - lw % vws
        nh, nw = (lh + pad_h) // vws, (lw + pad_w) // vws
        ip = F.pad(index, (0, pad_w, 0, pad_h), "constant", -100)
        ip = ip.reshape(t, nh, vws, nw, vws).permute(0, 1, 3, 2, 4).reshape(t, nh * nw, vws, vws)
        seqlens = (ip != -100).sum([2, 3]).reshape(-1)
        ip = ip.reshape(-1)
        window_index.append(ip[ip != -100] + base)
        cu.extend((seqlens.cumsum(0) * unit + cu[-1]).tolist())
        base += t * lh * lw
    cu = torch.unique_consecutive(torch.tensor(cu, dtype=torch.int64))
    return torch.cat(window_index), cu


def _segment_attention(q, k, v, cu, scale):
    """Non-causal attention inside each [cu[i], cu[i+1]) segment. q,k,v: [S, H, d]."""
    out = torch.empty_like(q)
    for a, b in zip(cu[:-1].tolist(), cu[1:].tolist()):
        s = torch.einsum("ihd,jhd->hij", q[a:b], k[a:b]) * scale
        out[a:b] = torch.einsum("hij,jhd->ihd", torch.softmax(s, -1), v[a:b])
    return out


def vit_forward(w, cfg, pixel_values, grid_thw):
    """pixel_values [N, C*T*P*P] fp32, grid_thw list of (t,h,w) -> merged embeds [N/4, hidden_size]."""
    q25 = cfg["family"] == "qwen2_5_vl"
    vd, H = cfg["vit_hidden"], cfg["vit_heads"]
    hd = vd // H
    unit = cfg["spatial_merge_size"] ** 2
    x = pixel_values.float() @ w["visual.patch_embed.proj.weight"].reshape(vd, -1).t()  # Conv3d == GEMM (Q25:106-114)
    N = x.shape[0]
    pos = vit_rot_pos(cfg, grid_thw)
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd // 2, 2, dtype=torch.float32) / (hd // 2)))  # Q25:124
    freqs = torch.cat([pos[:, 0:1].float() * inv, pos[:, 1:2].float() * inv], -1)           # [N, hd/2]
    cu_full = torch.tensor([0] + [h * w for t, h, w in grid_thw for _ in range(t)]).cumsum(0)
    if q25:
        widx, cu_win = vit_window_index(cfg, grid_thw)
        x = x.reshape(N // unit, unit, -1)[widx].reshape(N, -1)
        freqs = freqs.reshape(N // unit, unit, -1)[widx].reshape(N, -1)
    emb = torch.cat([freqs, freqs], -1)
    cos, sin = emb.cos()[:, None, :], emb.sin()[:, None, :]

    def rot(t):  # Q25:149-167
        h2 = t.shape[-1] // 2
        return t * cos + torch.cat([-t[..., h2:], t[..., :h2]], -1) * sin

    for i in range(cfg["vit_depth"]):
        p = f"visual.blocks.{i}."
        if q25:
            h1 = rms_norm(x, w[p + "norm1.weight"], 1e-6)
        else:
            h1 = F.layer_norm(x, (vd,), w[p + "norm1.weight"], w[p + "norm1.bias"], 1e-6)
        qkv = (h1 @ w[p + "attn.qkv.weight"].t() + w[p + "attn.qkv.bias"]).reshape(N, 3, H, hd)
        q, k, v = rot(qkv[:, 0]), rot(qkv[:, 1]), qkv[:, 2]
        cu = cu_full if (not q25 or i in cfg["fullatt_block_indexes"]) else cu_win
        a = _segment_attention(q, k, v, cu, hd ** -0.5).reshape(N, vd)
        x = x + a @ w[p + "attn.proj.weight"].t() + w[p + "attn.proj.bias"]
        if q25:
            h2 = rms_norm(x, w[p + "norm2.weight"], 1e-6)
            g = h2 @ w[p + "mlp.gate_proj.weight"].t() + w[p + "mlp.gate_proj.bias"]
            u = h2 @ w[p + "mlp.up_proj.weight"].t() + w[p + "mlp.up_proj.bias"]
            x = x + (F.silu(g) * u) @ w[p + "mlp.down_proj.weight"].t() + w[p + "mlp.down_proj.bias"]
        else:
            h2 = F.layer_norm(x, (vd,), w[p + "norm2.weight"], w[p + "norm2.bias"], 1e-6)
            f1 = h2 @ w[p + "mlp.fc1.weight"].t() + w[p + "mlp.fc1.bias"]
            x = x + (f1 * torch.sigmoid(1.702 * f1)) @ w[p + "mlp.fc2.weight"].t() + w[p + "mlp.fc2.bias"]
    if q25:
        y = rms_norm(x, w["visual.merger.ln_q.weight"], 1e-6)
    else:
        y = F.layer_norm(x, (vd,), w["visual.merger.ln_q.weight"], w["visual.merger.ln_q.bias"], 1e-6)
    y = y.reshape(N // unit, unit * vd)
    y = F.gelu(y @ w["visual.merger.mlp.0.weight"].t() + w["visual.merger.mlp.0.bias"])
    y = y @ w["visual.merger.mlp.2.weight"].t() + w["visual.merger.mlp.2.bias"]
    if q25:
        y = y[torch.argsort(widx)]
    return y


# ------------------------------------------------------------------------------------------------
# M-RoPE indices — transformers 4.51.1 semantics (pinned; see module docstring)
# ------------------------------------------------------------------------------------------------
def mrope_positions(cfg, input_ids, video_grid_thw, second_per_grid_ts=None):
    """input_ids: 1-D list/array for ONE un-padded sequence. Returns (pos [3, L] int64, rope_delta)."""
    ids = [int(t) for t in input_ids]
    ms = cfg["spatial_merge_size"]
    q25 = cfg["family"] == "qwen2_5_vl"
    pos_chunks, st, vid = [], 0, 0
    i, L = 0, len(ids)
    next_start = 0
    while i < L:
        if ids[i] == cfg["video_token_id"] or ids[i] == cfg["image_token_id"]:
            t, h, w = video_grid_thw[vid]
            lh, lw = h // ms, w // ms
            n = t * lh * lw
            spg = 1.0 if second_per_grid_ts is None else float(second_per_grid_ts[vid])
            tt = torch.arange(t).view(-1, 1).expand(-1, lh * lw)
            if q25:
                tt = (tt.float() * spg * cfg["tokens_per_second"]).long()
            tt = tt.flatten()
            hh = torch.arange(lh).view(1, -1, 1).expand(t, -1, lw).flatten()
            ww = torch.arange(lw).view(1, 1, -1).expand(t, lh, -1).flatten()
            pos_chunks.append(torch.stack([tt, hh, ww]) + next_start)
            next_start = int(pos_chunks[-1].max()) + 1
            i += n
            vid += 1
        else:
            j = i
            while j < L and ids[j] != cfg["video_token_id"] and ids[j] != cfg["image_token_id"]:
                j += 1
            n = j - i
            pos_chunks.append(torch.arange(n).view(1, -1).expand(3, -1) + next_start)
            next_start = int(pos_chunks[-1].max()) + 1
            i = j
    pos = torch.cat(pos_chunks, 1)
    delta = int(pos.max()) + 1 - L
    return pos, delta


# ------------------------------------------------------------------------------------------------
# LLM (Q25:860-942)
# ------------------------------------------------------------------------------------------------
def mrope_cos_sin(cfg, pos3):
    """pos3 [3, L] -> cos, sin [L, head_dim] with interleaved sections (Q25:595-608, :659-665)."""
    hd = cfg["head_dim"]
    inv = 1.0 / (cfg["rope_theta"] ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    freqs = pos3[:, :, None].float() * inv[None, None, :]          # [3, L, hd/2]
    emb = torch.cat([freqs, freqs], -1)                             # [3, L, hd]
    sec = list(cfg["mrope_section"]) * 2
    cos = torch.cat([c[i % 3] for i, c in enumerate(emb.cos().split(sec, -1))], -1)
    sin = torch.cat([s[i % 3] for i, s in enumerate(emb.sin().split(sec, -1))], -1)
    return cos, sin


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], -1)


def llm_forward(w, cfg, inputs_embeds, pos3, attn_bias=None, return_hidden=False, kv_cache=None):
    """Single sequence. inputs_embeds [L, D]; pos3 [3, L]; attn_bias [L, Ltot] additive (None = causal).
    kv_cache: optional list of (k, v) per layer to append to (decode). Returns logits [L, V] (or hidden)."""
    D, Hq, Hkv, hd = cfg["hidden_size"], cfg["num_attention_heads"], cfg["num_key_value_heads"], cfg["head_dim"]
    eps = cfg["rms_norm_eps"]
    x = inputs_embeds
    L = x.shape[0]
    cos, sin = mrope_cos_sin(cfg, pos3)
    cos, sin = cos[:, None, :], sin[:, None, :]
    for i in range(cfg["num_hidden_layers"]):
        p = f"model.layers.{i}."
        h = rms_norm(x, w[p + "input_layernorm.weight"], eps)
        q = (h @ w[p + "self_attn.q_proj.weight"].t() + w[p + "self_attn.q_proj.bias"]).view(L, Hq, hd)
        k = (h @ w[p + "self_attn.k_proj.weight"].t() + w[p + "self_attn.k_proj.bias"]).view(L, Hkv, hd)
        v = (h @ w[p + "self_attn.v_proj.weight"].t() + w[p + "self_attn.v_proj.bias"]).view(L, Hkv, hd)
        q = q * cos + _rot_half(q) * sin
        k = k * cos + _rot_half(k) * sin
        if kv_cache is not None:
            if kv_cache[i] is not None:
                k = torch.cat([kv_cache[i][0], k], 0)
                v = torch.cat([kv_cache[i][1], v], 0)
            kv_cache[i] = (k, v)
        Lk = k.shape[0]
        kk = k.repeat_interleave(Hq // Hkv, 1)
        vv = v.repeat_interleave(Hq // Hkv, 1)
        s = torch.einsum("ihd,jhd->hij", q, kk) * hd ** -0.5
        if attn_bias is None:
            ii = torch.arange(L)[:, None] + (Lk - L)
            s = s.masked_fill(torch.arange(Lk)[None, :] > ii, float("-inf"))
        else:
            s = s + attn_bias[None]
        a = torch.einsum("hij,jhd->ihd", torch.softmax(s, -1), vv).reshape(L, Hq * hd)
        x = x + a @ w[p + "self_attn.o_proj.weight"].t()
        h = rms_norm(x, w[p + "post_attention_layernorm.weight"], eps)
        g = F.silu(h @ w[p + "mlp.gate_proj.weight"].t()) * (h @ w[p + "mlp.up_proj.weight"].t())
        x = x + g @ w[p + "mlp.down_proj.weight"].t()
    x = rms_norm(x, w["model.norm.weight"], eps)
    if return_hidden:
        return x
    head = w["model.embed_tokens.weight"] if cfg["tie_word_embeddings"] else w["lm_head.weight"]
    return x @ head.t()


def embed_with_vision(w, cfg, input_ids, vision_embeds):
    """embed_tokens + masked_scatter of the video embeddings at video_token_id positions (Q25:1298-1315)."""
    ids = torch.as_tensor(input_ids, dtype=torch.long)
    x = w["model.embed_tokens.weight"][ids].clone()
    m = ids == cfg["video_token_id"]
    assert int(m.sum()) == vision_embeds.shape[0], "video token / feature count mismatch"
    x[m] = vision_embeds
    return x


def full_forward_logits(w, cfg, input_ids, pixel_values, grid_thw, second_per_grid_ts=None):
    """Logits [L, V] for one sequence (prompt incl. video tokens [+ completion])."""
    ve = vit_forward(w, cfg, pixel_values, grid_thw)
    x = embed_with_vision(w, cfg, input_ids, ve)
    pos, _ = mrope_positions(cfg, input_ids, grid_thw, second_per_grid_ts)
    return llm_forward(w, cfg, x, pos)


# ------------------------------------------------------------------------------------------------
# scoring + GRPO numerics (timer1_trainer.py:449-481, 581-590, 635-639, 700-777)
# ------------------------------------------------------------------------------------------------
def per_token_logps(logits, input_ids):
    """timer1_trainer.py:458-481 for one sequence: logits [L, V], ids [L] -> (logp, entropy) each [L-1]."""
    lg = logits[:-1]
    ids = torch.as_tensor(input_ids, dtype=torch.long)[1:]
    lsm = lg.log_softmax(-1)
    lp = lsm.gather(1, ids[:, None])[:, 0]
    ent = -(lsm.exp() * lsm).sum(-1)
    return lp, ent


def completion_mask(completion_ids, eos_token_id):
    """timer1_trainer.py:581-590."""
    is_eos = completion_ids == eos_token_id
    eos_idx = torch.full((is_eos.size(0),), is_eos.size(1), dtype=torch.long)
    eos_idx[is_eos.any(1)] = is_eos.int().argmax(1)[is_eos.any(1)]
    return (torch.arange(is_eos.size(1)).expand(is_eos.size(0), -1) <= eos_idx[:, None]).int()


def grpo_loss(logps, ref_logps, mask, rewards_per_func, G, beta, eps_low=0.2, eps_high=0.2, use_grpo=False,
              entropy=None):
    """timer1_trainer.py:635-639,700-777. logps may require grad. Returns (loss, metrics dict, advantages)."""
    rewards = rewards_per_func.sum(1)
    mean_g = rewards.view(-1, G).mean(1).repeat_interleave(G)
    std_g = rewards.view(-1, G).std(1).repeat_interleave(G)
    adv = (rewards - mean_g) / (std_g + 1e-4)
    kl = None
    if beta != 0.0:
        kl = torch.exp(ref_logps - logps) - (ref_logps - logps) - 1
    if use_grpo:
        ptl = torch.exp(logps - logps.detach()) * adv[:, None]
        ptl = -(ptl - beta * kl) if beta != 0.0 else -ptl
        loss = ((ptl * mask).sum(1) / mask.sum(1)).mean()
    else:
        c1 = torch.exp(logps - logps.detach())
        c2 = torch.clamp(c1, 1 - eps_low, 1 + eps_high)
        ptl = -torch.min(c1 * adv[:, None], c2 * adv[:, None])
        if beta != 0.0:
            ptl = ptl + beta * kl
        loss = (ptl * mask).sum() / mask.sum()
    metrics = {
        "completion_length": mask.sum(1).float().mean().item(),
        "reward": rewards.mean().item(),
        "reward_std": std_g.mean().item(),
    }
    if kl is not None:
        metrics["kl"] = ((kl * mask).sum(1) / mask.sum(1)).mean().item()
    if entropy is not None:
        metrics["generation_entropy"] = ((entropy * mask).sum(1) / mask.sum(1).clamp(min=1)).mean().item()
    return loss, metrics, adv


# ------------------------------------------------------------------------------------------------
# sampling (HF generate defaults of 4.51.1: temperature -> top-k 50 -> softmax -> multinomial)
# ------------------------------------------------------------------------------------------------
def topk_filtered_probs(logits, temperature, top_k):
    s = logits.float() / temperature
    if top_k and top_k < s.shape[-1]:
        kth = torch.topk(s, top_k)[0][..., -1, None]
        s = s.masked_fill(s < kth, float("-inf"))
    return torch.softmax(s, -1)


def generate(w, cfg, input_ids, pixel_values, grid_thw, G, max_new_tokens, temperature=1.0, top_k=50, seed=42,
             second_per_grid_ts=None, greedy=False, share_prefix=False):
    """Reference-style rollout: G independent copies (ViT + prefill per copy unless share_prefix) + KV-cached
    decode. Returns completion ids [G, max_new_tokens] (pad after eos) — timer1_trainer.py:568-577."""
    gen = torch.Generator().manual_seed(seed)
    eos = {cfg["eos_token_id"], cfg["pad_token_id"]}
    P = len(input_ids)
    out = torch.full((G, max_new_tokens), cfg["pad_token_id"], dtype=torch.long)
    shared = None
    for g in range(G):
        if shared is None or not share_prefix:
            ve = vit_forward(w, cfg, pixel_values, grid_thw)
            x = embed_with_vision(w, cfg, input_ids, ve)
            pos, delta = mrope_positions(cfg, input_ids, grid_thw, second_per_grid_ts)
            cache = [None] * cfg["num_hidden_layers"]
            logits = llm_forward(w, cfg, x, pos, kv_cache=cache)[-1]
            shared = (logits, [(k.clone(), v.clone()) for k, v in cache], delta)
        logits, cache, delta = shared[0], [(k.clone(), v.clone()) for k, v in shared[1]], shared[2]
        for t in range(max_new_tokens):
            if greedy:
                tok = int(logits.argmax())
            else:
                tok = int(torch.multinomial(topk_filtered_probs(logits, temperature, top_k), 1, generator=gen))
            out[g, t] = tok
            if tok in eos:
                break
            if t + 1 < max_new_tokens:
                p = torch.full((3, 1), P + t + delta, dtype=torch.long)
                logits = llm_forward(w, cfg, w["model.embed_tokens.weight"][tok][None], p, kv_cache=cache)[-1]
    return out


def score_completions(w, cfg, prompt_ids, completion_ids, pixel_values, grid_thw, second_per_grid_ts=None, tail_only=False):
    """Teacher-forced per-token logp / entropy of each completion, reference layout: returns [G, C] each
    (position P-1+c predicts completion token c).  timer1_trainer.py:601-612.
    tail_only: apply lm_head to the rows whose logits are returned only (identical values; the reference computes and
    discards the prompt rows' logits, which is what the CPU baseline times with tail_only=False)."""
    P = len(prompt_ids)
    ve = vit_forward(w, cfg, pixel_values, grid_thw)
    head = w["model.embed_tokens.weight"] if cfg["tie_word_embeddings"] else w["lm_head.weight"]
    lps, ents = [], []
    for row in completion_ids:
        ids = list(prompt_ids) + [int(t) for t in row]
        x = embed_with_vision(w, cfg, ids, ve)
        pos, _ = mrope_positions(cfg, ids, grid_thw, second_per_grid_ts)
        if tail_only:
            hN = llm_forward(w, cfg, x, pos, return_hidden=True)
            lp, ent = per_token_logps(hN[P - 1:] @ head.t(), ids[P - 1:])
        else:
            lp, ent = per_token_logps(llm_forward(w, cfg, x, pos), ids)
            lp, ent = lp[P - 1:], ent[P - 1:]
        lps.append(lp)
        ents.append(ent)
    return torch.stack(lps), torch.stack(ents)
