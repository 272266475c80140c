"""ORACLE tooling — what "parity" can mean for a bf16 model: the error of the REFERENCE'S OWN arithmetic.

The reference runs HF Qwen2.5-VL in bf16 (timer1_trainer.py:244-251: torch_dtype=torch.bfloat16) and takes log-softmax /
entropy on bf16 logits (:458-481).  north_star asks for 1e-3 relative against "the reference's own rollout"; against an
fp32 oracle no bf16 implementation — the reference included — gets there.  This script measures how far the installed HF
classes in bf16 (eager attention) land from the fp32 oracle on the SAME weights and inputs the GPU tests use, and writes
the errors to tests/golden/hf_bf16_error_bar.json; tests/test_engine_gpu.py and tests/test_parity_7b_width_gpu.py then
require the CUDA engine's error against the fp32 oracle to be <= 1.5 x the reference's (and never above the old fixed bars).

Cases: "toy" (config.toy, the inputs of test_grpo_step_loss_and_gradients_match_oracle) and "7b_d2" (Qwen2.5-VL-7B widths
— D 3584, I 18944, 28q/4kv x 128, V 152064, ViT 1280/3420/16 heads — at depth 2 + 2, grid (2,32,32), G 8, C 16).

Build container only (about 5 minutes, 25 GB of host memory):   python oracle/gen_golden_bf16_bar.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hf_reference as H, qwen_vl_oracle as O  # noqa: E402
from oracle.gen_golden_model import load_weights  # noqa: E402
from time_r1_b200 import config  # noqa: E402

WATCH = ["model.layers.0.self_attn.q_proj.weight", "model.layers.0.self_attn.k_proj.bias", "model.layers.1.mlp.gate_proj.weight",
         "model.layers.1.mlp.down_proj.weight", "model.layers.0.input_layernorm.weight", "model.norm.weight",
         "lm_head.weight", "model.embed_tokens.weight", "visual.merger.mlp.0.weight", "visual.merger.mlp.2.bias",
         "visual.merger.ln_q.weight", "model.layers.1.self_attn.o_proj.weight", "model.layers.1.self_attn.v_proj.weight"]
PERTURB = {"model.layers.1.mlp.down_proj.weight": 1.5, "visual.merger.mlp.2.weight": 1.2}


def cfg_7b_d2():
    import dataclasses

    return dataclasses.replace(config.PRESETS["qwen2.5-vl-7b"](), num_hidden_layers=2, vit_depth=2, fullatt_block_indexes=(1,))


def torch_seeded_weights(ocfg, seed, std=0.02):
    """bf16-representable fp32 weights from torch's CPU generator (numpy is too slow for 152064 x 3584 matrices); norm weights
    1 + 0.1 N(0,1).  Deterministic for a given torch build — the GPU box runs this same image."""
    from oracle.cpu_baseline import _shapes

    g = torch.Generator().manual_seed(seed)
    q25 = ocfg["family"] == "qwen2_5_vl"
    w = {}
    for n, shp in _shapes(ocfg).items():
        if q25 and (".mlp.fc" in n or n.endswith("norm1.bias") or n.endswith("norm2.bias") or n.endswith("ln_q.bias")):
            continue
        if not q25 and n.startswith("visual.blocks") and ("gate_proj" in n or "up_proj" in n or "down_proj" in n):
            continue
        if ocfg["tie_word_embeddings"] and n == "lm_head.weight":
            continue
        is_norm_w = n.endswith(("norm1.weight", "norm2.weight", "layernorm.weight", "model.norm.weight", "ln_q.weight"))
        t = 1.0 + 0.1 * torch.randn(shp, generator=g) if is_norm_w else torch.randn(shp, generator=g) * std
        w[n] = t.to(torch.bfloat16).float()
    return w


def case_inputs(name):
    """(cfg, ocfg, w_ref (fp32, bf16-representable), w_pol, pix, ids, grid, comp, rewards, beta) — shared with the GPU tests."""
    if name == "toy":
        cfg = config.toy("qwen2_5_vl")
        ocfg = O.default_cfg(**{k: v for k, v in cfg.to_dict().items() if k in O.default_cfg()})
        w = O.seeded_weights(ocfg, seed=7, std=0.05)
        rng = np.random.default_rng(3)
        grid, G, C = [(2, 8, 12)], 4, 12
        N = 2 * 8 * 12
        pix = torch.from_numpy(rng.standard_normal((N, cfg.patch_dim), dtype=np.float32)).to(torch.bfloat16).float()
        ids = np.array([5, 6, cfg.vision_start_token_id] + [cfg.video_token_id] * (N // 4) + [9] + list(rng.integers(10, 2000, 11)) + [7, 8])
        rng = np.random.default_rng(9)
        comp = torch.from_numpy(rng.integers(10, 2000, (G, C))).int()
        comp[1, 5] = cfg.eos_token_id
        comp[1, 6:] = cfg.pad_token_id
        rewards = rng.random((G, 2)).astype(np.float32)
    else:
        cfg = cfg_7b_d2()
        ocfg = O.default_cfg(**{k: v for k, v in cfg.to_dict().items() if k in O.default_cfg()})
        w = torch_seeded_weights(ocfg, seed=21)
        rng = np.random.default_rng(4)
        grid, G, C = [(2, 32, 32)], 8, 16
        N = 2 * 32 * 32
        pix = torch.from_numpy(rng.standard_normal((N, cfg.patch_dim), dtype=np.float32)).to(torch.bfloat16).float()
        ids = np.array([151644, 872, 198, cfg.vision_start_token_id] + [cfg.video_token_id] * (N // 4) + [151653] + list(rng.integers(0, 151643, 18))
                       + [151645, 198, 151644, 77091, 198])
        comp = torch.from_numpy(rng.integers(0, 151643, (G, C))).int()
        comp[2, 9] = cfg.eos_token_id
        comp[2, 10:] = cfg.pad_token_id
        comp[5, 0] = cfg.eos_token_id
        comp[5, 1:] = cfg.pad_token_id
        rewards = rng.random((G, 2)).astype(np.float32)
    w_pol = dict(w)
    for k, f in PERTURB.items():                      # move the policy away from the reference model: non-trivial KL
        w_pol[k] = (w[k].to(torch.bfloat16) * f).float()
    return cfg, ocfg, w, w_pol, pix, ids, grid, comp, rewards, 0.04


def oracle_step(ocfg, cfg, w, w_pol, pix, ids, grid, comp, rewards, beta, use_grpo=False):
    w_pol = {k: v.clone() for k, v in w_pol.items()}
    for k in WATCH:
        w_pol[k].requires_grad_(True)
    G = comp.shape[0]
    lp, ent = O.score_completions(w_pol, ocfg, list(ids), comp.numpy(), pix, grid, tail_only=True)
    with torch.no_grad():
        lp_ref, _ = O.score_completions(w, ocfg, list(ids), comp.numpy(), pix, grid, tail_only=True)
    mask = O.completion_mask(comp.long(), cfg.eos_token_id)
    loss, metrics, adv = O.grpo_loss(lp, lp_ref, mask, torch.from_numpy(rewards), G, beta, use_grpo=use_grpo, entropy=ent.detach())
    loss.backward()
    return dict(lp=lp.detach(), ent=ent.detach(), lp_ref=lp_ref, loss=loss.detach(), metrics=metrics, adv=adv, mask=mask,
                grads={k: w_pol[k].grad for k in WATCH})


def hf_bf16_step(cfg, ocfg, w, w_pol, pix, ids, grid, comp, rewards, beta, use_grpo=False):
    """The reference's arithmetic: HF classes in bf16, logits at every position, bf16 log-softmax (timer1_trainer.py:449-481)."""
    G, C = comp.shape
    P = len(ids)

    def model_of(wd):
        m = H.hf_model_class(cfg)(H.hf_config(cfg, "eager")).float().eval()
        load_weights(m, wd, cfg)
        return m.to(torch.bfloat16)

    pol, ref = model_of(w_pol), model_of(w)
    for p_ in ref.parameters():
        p_.requires_grad_(False)
    pc = torch.cat([torch.from_numpy(np.asarray(ids)).long()[None].repeat(G, 1), comp.long()], 1)
    pos, _ = O.mrope_positions(ocfg, list(ids) + [0] * C, grid, None)            # scoring pass: no second_per_grid_ts (:601-607)
    position_ids = pos[:, None, :].repeat(1, G, 1)
    mask = H.completion_mask_of(comp.long(), cfg.eos_token_id)
    attn = torch.cat([torch.ones(G, P, dtype=torch.long), mask.long()], 1)
    pixb = pix.to(torch.bfloat16).repeat(G, 1)
    gridt = torch.tensor(grid).repeat_interleave(G, 0)

    def logps(m):
        logits = m(input_ids=pc, attention_mask=attn, pixel_values_videos=pixb, video_grid_thw=gridt, position_ids=position_ids).logits
        logits = logits[:, :-1]
        tgt = pc[:, 1:]
        lps, ents = [], []
        for lr, ir in zip(logits, tgt):
            lsm = lr.log_softmax(-1)
            lps.append(torch.gather(lsm, 1, ir[:, None])[:, 0])
            pr = torch.exp(lsm)
            ents.append(-(pr * lsm).sum(-1))
        return torch.stack(lps)[:, P - 1:], torch.stack(ents)[:, P - 1:]

    lp, ent = logps(pol)
    with torch.no_grad():
        lp_ref, _ = logps(ref)
    loss, adv, kl = H.grpo_loss_of(lp, lp_ref, mask, torch.from_numpy(rewards), G, beta, use_grpo=use_grpo)
    loss.backward()
    sd = dict(pol.named_parameters())
    grads = {}
    for k in WATCH:
        hk = "model." + k if k.startswith("visual.") else ("model.language_model." + k[len("model."):] if k.startswith("model.") else k)
        grads[k] = sd[hk].grad.float()
    return dict(lp=lp.detach().float(), ent=ent.detach().float(), lp_ref=lp_ref.float(), loss=loss.detach().float(), grads=grads, mask=mask)


def errors(o, h):
    m = o["mask"].bool()
    rec = dict(logp_max_abs=float((h["lp"] - o["lp"]).abs()[m].max()), logp_all_max_abs=float((h["lp"] - o["lp"]).abs().max()),
               entropy_max_abs=float((h["ent"] - o["ent"]).abs().max()), ref_logp_max_abs=float((h["lp_ref"] - o["lp_ref"]).abs().max()),
               loss_abs=float((h["loss"] - o["loss"]).abs()), loss_oracle=float(o["loss"]), grads={})
    for k in WATCH:
        g, r = h["grads"][k].reshape(-1), o["grads"][k].reshape(-1)
        rec["grads"][k] = dict(max_rel=float((g - r).abs().max() / (r.abs().max() + 1e-12)),
                               cosine=float(torch.nn.functional.cosine_similarity(g.double(), r.double(), dim=0)), oracle_absmax=float(r.abs().max()))
    return rec


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    out = {"how": "HF classes (transformers %s) bf16 eager vs oracle/qwen_vl_oracle.py fp32 on identical bf16-representable weights and "
                  "inputs; teacher-forced scoring of fixed completions, beta 0.04, default (PPO-clip) loss" % __import__("transformers").__version__,
           "torch": torch.__version__}
    for name in ("toy", "7b_d2"):
        t0 = time.time()
        cfg, ocfg, w, w_pol, pix, ids, grid, comp, rewards, beta = case_inputs(name)
        o = oracle_step(ocfg, cfg, w, w_pol, pix, ids, grid, comp, rewards, beta)
        t1 = time.time()
        h = hf_bf16_step(cfg, ocfg, w, w_pol, pix, ids, grid, comp, rewards, beta)
        rec = errors(o, h)
        rec.update(oracle_s=round(t1 - t0, 1), hf_bf16_s=round(time.time() - t1, 1), P=len(ids), G=int(comp.shape[0]), C=int(comp.shape[1]),
                   grid=list(grid[0]), oracle_logp_checksum=float(o["lp"].double().sum()), oracle_kl=o["metrics"]["kl"])
        out[name] = rec
        print(name, json.dumps({k: v for k, v in rec.items() if k != "grads"}))
        for k, v in rec["grads"].items():
            print("   ", k, v)
    with open(os.path.join(ROOT, "tests", "golden", "hf_bf16_error_bar.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
