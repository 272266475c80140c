"""CPU suite, part 1: the oracle is pinned to the reference's own known answers before anything trusts it.

  * tests/golden/iou_train_2k5.json     2500 rows of the reference's annotation file: difficulty == 100 * IoU(pred, gt)
  * tests/golden/rewards_golden.json    outputs of the UNMODIFIED reference reward functions (oracle/gen_golden_rewards.py)
  * tests/golden/hf_toy_*.npz           HF Qwen2.5-VL / Qwen2-VL outputs on a seeded toy config (oracle/gen_golden_model.py)
  * tests/golden/mrope_vllm.json        vLLM's 4.51.1-semantics M-RoPE indices (oracle/gen_golden_mrope.py)
"""
import ctypes
import json
import math
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _c_oracle():
    so = os.path.join(ROOT, "oracle", "libgrpo_oracle.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)
    lib = ctypes.CDLL(so)
    lib.oracle_iou.restype = ctypes.c_double
    lib.oracle_iou.argtypes = [ctypes.c_int] + [ctypes.c_double] * 5
    return lib


def test_iou_known_answers_2500_rows():
    from oracle import rewards_oracle as R

    rows = json.load(open(os.path.join(GOLD, "iou_train_2k5.json")))
    assert len(rows) == 2500
    c = _c_oracle()
    for r in rows:
        assert R.difficulty(r["pred"], r["gt"]) == r["difficulty"]                      # exact, python restatement
        v = c.oracle_iou(0, r["pred"][0], r["pred"][1], r["gt"][0], r["gt"][1], r["duration"])
        assert (0.0 if math.isnan(v) else v * 100) == r["difficulty"]                   # exact, C restatement


def test_reward_functions_match_reference_outputs():
    from oracle import rewards_oracle as R
    from time_r1_b200 import rewards as P

    rows = json.load(open(os.path.join(GOLD, "rewards_golden.json")))
    c = _c_oracle()
    unbound = 0
    for r in rows:
        parsed = R.parse_timestamp_output(r["text"])
        assert (list(parsed) if parsed else None) == r["parsed"]
        # the product's host-side parser / format check are the same spec
        pp = P.parse_timestamp_output(r["text"])
        assert (list(pp) if pp else None) == r["parsed"]
        assert R.format_reward(r["text"]) == r["format"] == P.format_reward([r["text"]])[0]
        assert R.iou_reward(r["text"], r["solution"]) == r["iou"]
        assert R.iou_v2_reward(r["text"], r["solution"], r["duration"]) == r["iou_v2"]
        if r["iou"] is None:
            unbound += 1
        elif parsed:
            for mode, key in ((0, "iou"), (1, "iou_v2")):
                assert c.oracle_iou(mode, parsed[0], parsed[1], r["solution"][0], r["solution"][1], r["duration"]) == r[key]
    assert unbound > 0, "the degenerate union<=0 rows must be present in the golden set"


@pytest.mark.parametrize("family", ["qwen2_5_vl", "qwen2_vl"])
def test_model_oracle_matches_hf_golden(family):
    from oracle import qwen_vl_oracle as O
    from time_r1_b200 import config

    g = np.load(os.path.join(GOLD, f"hf_toy_{family}.npz"))
    cfg = config.toy(family)
    ocfg = O.default_cfg(**{k: v for k, v in cfg.to_dict().items() if k in O.default_cfg()})
    w = O.seeded_weights(ocfg, seed=7, std=0.05)
    rng = np.random.default_rng(3)
    grid = [tuple(int(v) for v in r) for r in g["grid"]]
    N = sum(t * h * w_ for t, h, w_ in grid)
    pix = torch.from_numpy(rng.standard_normal((N, cfg.patch_dim), dtype=np.float32))
    ids = [int(t) for t in g["ids"]]
    emb = O.vit_forward(w, ocfg, pix, grid)
    assert np.abs(emb.numpy() - g["vision_embeds"]).max() < 5e-5          # fp32 vs fp32, different summation order
    logits = O.full_forward_logits(w, ocfg, ids, pix, grid, [1.0])
    assert np.abs(logits[-8:].numpy() - g["logits_tail"]).max() < 5e-5
    pos, _ = O.mrope_positions(ocfg, ids, grid, [1.0])
    assert np.array_equal(pos.numpy(), g["pos"])


def test_mrope_indices_match_vllm_4_51_semantics():
    from oracle import qwen_vl_oracle as O
    from time_r1_b200 import config, indexing

    cases = json.load(open(os.path.join(GOLD, "mrope_vllm.json")))
    assert len(cases) >= 6
    for c in cases:
        ocfg = O.default_cfg(tokens_per_second=c["tokens_per_second"])
        pos, delta = O.mrope_positions(ocfg, c["ids"], [tuple(g) for g in c["grids"]], c["second_per_grid_ts"])
        assert pos.tolist() == c["pos"] and delta == c["delta"]
        pcfg = config.ModelConfig(tokens_per_second=c["tokens_per_second"])
        p2, d2 = indexing.mrope_positions(pcfg, c["ids"], c["grids"], c["second_per_grid_ts"])   # product host logic
        assert p2.tolist() == c["pos"] and d2 == c["delta"]


def test_grpo_loss_python_vs_c_restatement_and_autograd():
    from oracle import qwen_vl_oracle as O

    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "libgrpo_oracle.so"))
    rng = np.random.default_rng(1)
    G, C = 8, 21
    logp = -rng.random((G, C)).astype(np.float32) * 3
    ref = (logp + 0.3 * rng.standard_normal((G, C))).astype(np.float32)
    rpf = rng.random((G, 2)).astype(np.float32)
    lens = rng.integers(1, C + 1, G)
    mask = (np.arange(C)[None, :] < lens[:, None]).astype(np.int32)
    for use_grpo in (0, 1):
        for beta in (0.0, 0.04):
            lp = torch.from_numpy(logp).clone().requires_grad_(True)
            loss, metrics, adv = O.grpo_loss(lp, torch.from_numpy(ref), torch.from_numpy(mask), torch.from_numpy(rpf), G, beta, use_grpo=bool(use_grpo))
            loss.backward()
            rew, cadv, mean, std = (np.zeros(G, np.float32), np.zeros(G, np.float32), np.zeros(1, np.float32), np.zeros(1, np.float32))
            f = lambda a: a.ctypes.data_as(ctypes.c_void_p)
            lib.oracle_group_adv(f(rpf), G, 2, G, f(rew), f(cadv), f(mean), f(std))
            assert np.allclose(cadv, adv.numpy(), atol=1e-5)
            out = np.zeros(3)
            dl = np.zeros((G, C))
            lib.oracle_grpo_loss(f(logp), f(ref), f(cadv), f(mask), G, C, ctypes.c_float(beta), use_grpo, f(out), f(dl))
            assert abs(out[0] - loss.item()) < 1e-5
            assert np.allclose(dl, lp.grad.numpy(), atol=1e-6)
            if beta:
                assert abs(out[1] - metrics["kl"]) < 1e-5
            assert abs(out[2] - metrics["completion_length"]) < 1e-6


def test_completion_mask_and_sampler_oracle_properties():
    from oracle import qwen_vl_oracle as O

    ids = torch.tensor([[5, 9, 7, 9, 1], [1, 2, 3, 4, 5], [9, 9, 9, 9, 9]])
    m = O.completion_mask(ids, 9)
    assert m.tolist() == [[1, 1, 0, 0, 0], [1, 1, 1, 1, 1], [1, 0, 0, 0, 0]]
    logits = torch.randn(3, 1000, generator=torch.Generator().manual_seed(0))
    p = O.topk_filtered_probs(logits, 0.7, 50)
    assert ((p > 0).sum(-1) == 50).all() and torch.allclose(p.sum(-1), torch.ones(3))
    assert torch.equal(p.argmax(-1), logits.argmax(-1))
