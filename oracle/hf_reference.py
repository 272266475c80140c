"""ORACLE / BASELINE tooling — the reference's own GRPO step driven through the HF model classes, for timing on the
GPU box (bench.py `--impl hf_gpu`, the `gpu_reference` leg of the BENCH line) and on the host cores (cfg1, CPU).

Only tests/, bench.py's baseline legs and the golden generators may import this module; the product never does.

`reference_style_step` restates `TimeR1_Trainer.compute_loss` + the HF Trainer step around it
(/root/reference/src/time_r1/rl/timer1_trainer.py, cited per block below) with REFERENCE-STYLE ARITHMETIC — nothing of
this repository's restructuring: HF `generate` on G expanded copies including the pixels (:568-573), two full-length
scoring passes on G-times-repeated pixels with logits at EVERY position and the per-row log-softmax loop (:449-481,
:594-632), rewards / advantage / clipped loss (:685-737), autograd backward through gradient-checkpointed layers
(`--gradient_checkpointing true`, scripts/posttrain/train_rl.sh:32), clip 1.0 and one AdamW step (HF Trainer defaults).
`oracle/gen_golden_compute_loss.py` checks, in the build container, that this restatement returns the same loss as the
UNMODIFIED `compute_loss` (imported from /root/reference with its missing dependencies stubbed) on the same model,
inputs and RNG state; /root/reference does not exist on the GPU box, so the restatement is what travels.

Favourable-to-the-reference choices (stated in the bench line): generation runs in eval mode WITH the KV cache (under
`--gradient_checkpointing true` the reference's model is in train mode during `generate`, where transformers drops the
cache in every checkpointed layer); no DeepSpeed ZeRO-3 parameter gathering and no CPU-offloaded optimizer
(scripts/zero3_offload.json) — weights, gradients and AdamW state stay in HBM; fused AdamW; flash-attention-2 when it
loads, else SDPA; no `torch.cuda.empty_cache()` (:779-780), no metric gathers / `.item()` syncs (:740-777).
"""
from __future__ import annotations

import copy
import time

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------
# model construction (random init, HF classes of the transformers version installed here)
# ------------------------------------------------------------------------------------------------
def hf_config(cfg, attn_impl: str = "eager"):
    """HF config object for a time_r1_b200.config.ModelConfig (either family)."""
    if cfg.family == "qwen2_5_vl":
        from transformers import Qwen2_5_VLConfig as C
        vis = dict(depth=cfg.vit_depth, hidden_size=cfg.vit_hidden, intermediate_size=cfg.vit_intermediate, num_heads=cfg.vit_heads,
                   patch_size=cfg.patch_size, temporal_patch_size=cfg.temporal_patch_size, spatial_merge_size=cfg.spatial_merge_size,
                   window_size=cfg.window_size, fullatt_block_indexes=list(cfg.fullatt_block_indexes), out_hidden_size=cfg.hidden_size,
                   tokens_per_second=cfg.tokens_per_second, hidden_act="silu")
    else:
        from transformers import Qwen2VLConfig as C
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
    hc._attn_implementation = attn_impl
    return hc


def hf_model_class(cfg):
    if cfg.family == "qwen2_5_vl":
        from transformers import Qwen2_5_VLForConditionalGeneration as M
    else:
        from transformers import Qwen2VLForConditionalGeneration as M
    return M


def build_random_model(cfg, device, dtype, attn_impl: str = "eager", seed: int = 0):
    """Random-init N(0, initializer_range) model of the architecture the reference instantiates (timer1_trainer.py:244)."""
    torch.manual_seed(seed)
    M = hf_model_class(cfg)
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with torch.device(device):
            m = M(hf_config(cfg, attn_impl))
    finally:
        torch.set_default_dtype(old)
    return m.to(dtype)


def pick_attn_impl(device) -> str:
    """flash_attention_2 (what every reference script passes, train_rl.sh:33) when the package loads AND runs on this
    device, else sdpa."""
    if torch.device(device).type != "cuda":
        return "sdpa"
    try:
        from flash_attn import flash_attn_func

        q = torch.randn(1, 16, 2, 64, device=device, dtype=torch.bfloat16)
        flash_attn_func(q, q, q, causal=True)
        torch.cuda.synchronize()
        return "flash_attention_2"
    except Exception:
        return "sdpa"


def freeze_vit_except_merger(model):
    """--fix_vit true (timer1_trainer.py:272-280)."""
    vis = model.visual if hasattr(model, "visual") else model.model.visual
    vis.requires_grad_(False)
    vis.merger.requires_grad_(True)


def create_reference_model(model):
    """trl.models.create_reference_model (timer1_trainer.py:305): deepcopy, frozen, eval."""
    ref = copy.deepcopy(model)
    for p in ref.parameters():
        p.requires_grad_(False)
    return ref.eval()


# ------------------------------------------------------------------------------------------------
# the step
# ------------------------------------------------------------------------------------------------
def get_per_token_logps(model, input_ids, attention_mask, pixel_values_videos, video_grid_thw):
    """timer1_trainer.py:449-481 — logits at every position, per-row log-softmax / gather / entropy loop."""
    logits = model(input_ids, attention_mask=attention_mask, pixel_values_videos=pixel_values_videos,
                   video_grid_thw=video_grid_thw).logits
    logits = logits[:, :-1, :]
    input_ids = input_ids[:, 1:]
    lps, ents = [], []
    for logits_row, ids_row in zip(logits, input_ids):
        log_probs = logits_row.log_softmax(dim=-1)
        lps.append(torch.gather(log_probs, dim=1, index=ids_row.unsqueeze(1)).squeeze(1))
        probs = torch.exp(log_probs)
        ents.append(-torch.sum(probs * log_probs, dim=-1))
    return torch.stack(lps), torch.stack(ents)


def completion_mask_of(completion_ids, eos_token_id):
    """timer1_trainer.py:581-590."""
    is_eos = completion_ids == eos_token_id
    dev = completion_ids.device
    eos_idx = torch.full((is_eos.size(0),), is_eos.size(1), dtype=torch.long, device=dev)
    eos_idx[is_eos.any(dim=1)] = is_eos.int().argmax(dim=1)[is_eos.any(dim=1)]
    seq = torch.arange(is_eos.size(1), device=dev).expand(is_eos.size(0), -1)
    return (seq <= eos_idx.unsqueeze(1)).int()


def grpo_loss_of(per_token_logps, ref_per_token_logps, completion_mask, rewards_per_func, G, beta, use_grpo=False,
                 eps_low=0.2, eps_high=0.2):
    """timer1_trainer.py:635-639, 700-737."""
    per_token_kl = None
    if beta != 0.0:
        d = ref_per_token_logps - per_token_logps
        per_token_kl = torch.exp(d) - d - 1
    rewards = rewards_per_func.sum(dim=1)
    mean_g = rewards.view(-1, G).mean(dim=1).repeat_interleave(G, dim=0)
    std_g = rewards.view(-1, G).std(dim=1).repeat_interleave(G, dim=0)
    advantages = (rewards - mean_g) / (std_g + 1e-4)
    if use_grpo:
        ptl = torch.exp(per_token_logps - per_token_logps.detach()) * advantages.unsqueeze(1)
        ptl = -(ptl - beta * per_token_kl) if beta != 0.0 else -ptl
        loss = ((ptl * completion_mask).sum(dim=1) / completion_mask.sum(dim=1)).mean()
    else:
        c1 = torch.exp(per_token_logps - per_token_logps.detach())
        c2 = torch.clamp(c1, 1 - eps_low, 1 + eps_high)
        ptl = -torch.min(c1 * advantages.unsqueeze(1), c2 * advantages.unsqueeze(1))
        if beta != 0.0:
            ptl = ptl + beta * per_token_kl
        loss = (ptl * completion_mask).sum() / completion_mask.sum()
    return loss, advantages, per_token_kl


class ReferenceStyleGRPO:
    """Holds policy + reference model + AdamW the way the reference's Trainer does and runs reference-style steps."""

    def __init__(self, cfg, device, dtype, G: int, C: int, beta: float = 0.04, temperature: float = 1.0, top_k: int = 50,
                 lr: float = 1e-6, grad_ckpt: bool = True, attn_impl: str | None = None, seed: int = 0, model=None):
        from transformers import GenerationConfig

        self.cfg, self.device, self.G, self.C, self.beta = cfg, torch.device(device), G, C, beta
        self.attn_impl = attn_impl or pick_attn_impl(device)
        self.model = model if model is not None else build_random_model(cfg, device, dtype, self.attn_impl, seed)
        freeze_vit_except_merger(self.model)
        self.ref_model = create_reference_model(self.model) if beta != 0.0 else None
        if grad_ckpt:
            self.model.gradient_checkpointing_enable()
            self.model.config.use_cache = False                      # timer1_trainer.py:237-241
        self.trainable = [p for p in self.model.parameters() if p.requires_grad]
        fused = self.device.type == "cuda"
        self.opt = torch.optim.AdamW(self.trainable, lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, fused=fused)
        # timer1_trainer.py:371-377 (+ top_k 50 = the 4.51.1 GenerationConfig default the reference inherits; the installed
        # transformers dropped that default, so it is set here)
        self.generation_config = GenerationConfig(max_new_tokens=C, do_sample=True, temperature=temperature, top_k=top_k,
                                                  num_return_sequences=G, pad_token_id=cfg.pad_token_id,
                                                  eos_token_id=[cfg.eos_token_id, cfg.pad_token_id])

    def step(self, prompt_inputs: dict, reward_fn, grad_accum: int = 1, do_optimizer: bool = True, sync=None) -> dict:
        """One micro-step (+ optimizer step).  prompt_inputs: input_ids [1, P], attention_mask [1, P],
        pixel_values_videos [N, 1176], video_grid_thw [1, 3] (+ second_per_grid_ts) on self.device.
        reward_fn(completion_ids [G, C] tensor) -> rewards_per_func [G, n] float32 tensor on the device."""
        G = self.G
        sync = sync or (torch.cuda.synchronize if self.device.type == "cuda" else (lambda: None))
        t = {}
        sync(); t0 = time.perf_counter()
        # ---- generate (:568-577) — eval mode so that the KV cache is used (see module docstring)
        self.model.eval()
        with torch.no_grad():
            pc_ids = self.model.generate(**prompt_inputs, generation_config=self.generation_config, use_cache=True)
        self.model.train()
        P = prompt_inputs["input_ids"].size(1)
        completion_ids = pc_ids[:, P:]
        prompt_mask = prompt_inputs["attention_mask"].repeat_interleave(G, dim=0)
        sync(); t["generate"] = time.perf_counter() - t0; t0 = time.perf_counter()
        # ---- mask, repeat the pixels G times (:581-599)
        cmask = completion_mask_of(completion_ids, self.cfg.eos_token_id)
        attention_mask = torch.cat([prompt_mask, cmask], dim=1)
        pix = prompt_inputs["pixel_values_videos"].repeat(G, 1)
        grid = prompt_inputs["video_grid_thw"].repeat_interleave(G, dim=0)
        # ---- policy scoring with grad, reference scoring without (:601-639)
        lp, ent = get_per_token_logps(self.model, pc_ids, attention_mask, pix, grid)
        lp, ent = lp[:, P - 1:], ent[:, P - 1:]
        sync(); t["score_policy"] = time.perf_counter() - t0; t0 = time.perf_counter()
        rlp = None
        if self.beta != 0.0:
            with torch.inference_mode():
                rlp, _ = get_per_token_logps(self.ref_model, pc_ids, attention_mask, pix, grid)
            rlp = rlp[:, P - 1:]
        sync(); t["score_ref"] = time.perf_counter() - t0; t0 = time.perf_counter()
        # ---- rewards, advantage, loss (:642-737)
        rpf = reward_fn(completion_ids).to(self.device, torch.float32)
        loss, adv, kl = grpo_loss_of(lp, rlp, cmask, rpf, G, self.beta)
        # ---- HF Trainer.training_step: loss / accum, backward; then clip + AdamW (HF trainer.py:1927-1934, 2493)
        (loss / grad_accum).backward()
        sync(); t["backward"] = time.perf_counter() - t0; t0 = time.perf_counter()
        gnorm = None
        if do_optimizer:
            gnorm = torch.nn.utils.clip_grad_norm_(self.trainable, 1.0)
            self.opt.step()
            self.opt.zero_grad(set_to_none=True)
        sync(); t["optimizer"] = time.perf_counter() - t0
        return dict(loss=loss.detach(), advantages=adv, completion_ids=completion_ids, completion_mask=cmask, logp=lp.detach(),
                    ref_logp=rlp, entropy=ent.detach(), grad_norm=gnorm, times=t, prompt_len=P)


def synthetic_prompt_inputs(cfg, grid, device, dtype, seed: int = 42, n_text: int = 110):
    """SURVEY §8d synthetic inputs, as the HF processor would hand them to the model (timer1_trainer.py:547-565)."""
    rng = np.random.default_rng(seed)
    t, h, w = grid
    N = t * h * w
    nv = N // cfg.merge_unit
    V = cfg.vocab_size
    lo_special = min(cfg.video_token_id, cfg.image_token_id, cfg.vision_start_token_id, cfg.eos_token_id, cfg.pad_token_id)
    text = rng.integers(0, max(8, min(151643, lo_special)), n_text)
    head = [151644, 872, 198] if V > 151644 else [1, 2, 3]
    tail = [cfg.eos_token_id, 198 % V] + ([151644, 77091, 198] if V > 151644 else [1, 4, 3])
    vend = cfg.vision_start_token_id + 1 if cfg.vision_start_token_id + 1 < V else cfg.vision_start_token_id - 1
    ids = np.concatenate([head, [cfg.vision_start_token_id], np.full(nv, cfg.video_token_id), [vend], text, tail]).astype(np.int64)
    g = torch.Generator().manual_seed(seed)
    pix = torch.randn((N, cfg.patch_dim), generator=g, dtype=torch.float32)
    inputs = dict(input_ids=torch.from_numpy(ids)[None].to(device), attention_mask=torch.ones(1, len(ids), dtype=torch.long, device=device),
                  pixel_values_videos=pix.to(device=device, dtype=dtype), video_grid_thw=torch.tensor([[t, h, w]], device=device))
    return inputs
