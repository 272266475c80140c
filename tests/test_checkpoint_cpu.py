"""CPU suite: HF-format checkpoint export / import (main.py:589-623, train_rl_SF.sh:76-105).

  * the exported directory loads with the installed HF `from_pretrained` and produces the oracle's logits for the same weights;
  * the processor / tokenizer and generation_config.json travel with every checkpoint (evaluate.py / the next filtering round
    call AutoProcessor.from_pretrained(checkpoint));
  * resume loads INTO the live flat buffers (no second model in HBM) and continues the step / LR arithmetic of main.py:589-618.
"""
import collections
import json
import os

import numpy as np
import pytest
import torch

from test_trainer_host_cpu import StandInEngine, make_trainer, rows


def toy_weights(family="qwen2_5_vl", seed=7):
    from oracle import qwen_vl_oracle as O
    from time_r1_b200 import config, weights

    cfg = config.toy(family)
    ocfg = O.default_cfg(**{k: v for k, v in cfg.to_dict().items() if k in O.default_cfg()})
    w_or = O.seeded_weights(ocfg, seed=seed, std=0.05)
    mw = weights.ModelWeights(cfg, "cpu")
    mw.load_state_dict(w_or)
    return O, cfg, ocfg, w_or, mw


class SavingProcessor:
    eos_token_id = 2043

    def save_pretrained(self, d):
        with open(os.path.join(d, "preprocessor_config.json"), "w") as f:
            json.dump({"processor_class": "Qwen2_5_VLProcessor"}, f)
        with open(os.path.join(d, "tokenizer_config.json"), "w") as f:
            json.dump({"tokenizer_class": "Qwen2Tokenizer"}, f)


@pytest.mark.parametrize("family", ["qwen2_5_vl", "qwen2_vl"])
def test_exported_checkpoint_loads_with_hf_from_pretrained_and_matches_the_oracle(tmp_path, family):
    from time_r1_b200.trainer import save_checkpoint_dir
    from transformers import AutoConfig, AutoModelForImageTextToText

    O, cfg, ocfg, w_or, mw = toy_weights(family)
    d = str(tmp_path / "export")
    os.makedirs(d)
    save_checkpoint_dir(d, cfg, mw)
    assert {"config.json", "generation_config.json", "model.safetensors"} <= set(os.listdir(d))
    hc = AutoConfig.from_pretrained(d)
    assert hc.model_type == ("qwen2_5_vl" if family == "qwen2_5_vl" else "qwen2_vl")
    m, info = AutoModelForImageTextToText.from_pretrained(d, torch_dtype=torch.float32, attn_implementation="eager", output_loading_info=True)
    assert not info["missing_keys"] and not info["unexpected_keys"] and not info.get("mismatched_keys"), info
    m.eval()
    rng = np.random.default_rng(3)
    grid = [(2, 8, 12)]
    N = 2 * 8 * 12
    pix = torch.from_numpy(rng.standard_normal((N, cfg.patch_dim), dtype=np.float32))
    ids = [5, 6, cfg.vision_start_token_id] + [cfg.video_token_id] * (N // 4) + [9, 10, 11, 12, 7, 8]
    pos, _ = O.mrope_positions(ocfg, ids, grid, [1.0])
    with torch.no_grad():
        out = m(input_ids=torch.tensor([ids]), pixel_values_videos=pix, video_grid_thw=torch.tensor(grid), position_ids=pos[:, None, :],
                attention_mask=torch.ones(1, len(ids), dtype=torch.long)).logits[0]
    ref = O.full_forward_logits(w_or, ocfg, ids, pix, grid, [1.0])
    assert (out - ref).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item())


def test_checkpoint_carries_processor_and_resume_loads_in_place(tmp_path):
    from time_r1_b200 import rewards
    from time_r1_b200.trainer import GRPOArguments, TimeR1_Trainer

    O, cfg, ocfg, w_or, mw = toy_weights()
    args = GRPOArguments(output_dir=str(tmp_path), num_generations=2, max_completion_length=4, beta=0.0, gradient_accumulation_steps=2,
                         num_train_epochs=2, logging_steps=1, save_strategy="steps", save_steps=2, learning_rate=1e-4, fix_vit=True)
    tr = make_trainer(TimeR1_Trainer, args, [rewards.format_reward])
    tr.model_cfg, tr.weights = cfg, mw
    tr.processing_class = type("P", (SavingProcessor,), {"batch_decode": lambda self, ids, skip_special_tokens=True: ["<think>a</think><answer>1 to 2</answer>"] * len(ids)})()
    eng = tr.engine
    n = mw.train.numel
    eng.master, eng.adam_m, eng.adam_v = mw.train.buf.float().clone(), torch.zeros(n), torch.ones(n) * 0.5
    eng.opt_step, eng.shard_lo, eng.shard = 7, 0, n
    tr.train()
    assert tr.state.global_step == 6 and tr.state.max_steps == 6              # 6 rows / accum 2 * 2 epochs
    ck = os.path.join(str(tmp_path), "checkpoint-4")
    assert {"config.json", "generation_config.json", "model.safetensors", "trainer_state.json", "preprocessor_config.json",
            "tokenizer_config.json", "optimizer_rank0.pt"} <= set(os.listdir(ck))
    gc = json.load(open(os.path.join(ck, "generation_config.json")))
    assert gc["eos_token_id"] == [cfg.eos_token_id, cfg.pad_token_id]
    # ---- resume, the way main.py:589-618 does it
    saved = mw.train.buf.clone()
    mw.train.buf.zero_()
    mw.frozen.buf.zero_()
    ptr_before = (mw.train.buf.data_ptr(), mw.frozen.buf.data_ptr())
    tr2 = make_trainer(TimeR1_Trainer, args, [rewards.format_reward])
    tr2.model_cfg, tr2.weights, tr2.processing_class, tr2.engine = cfg, mw, tr.processing_class, StandInEngine(2, 4)
    e2 = tr2.engine
    e2.master, e2.adam_m, e2.adam_v, e2.opt_step, e2.shard_lo, e2.shard = torch.zeros(n), torch.zeros(n), torch.zeros(n), 0, 0, n
    resumed = json.load(open(os.path.join(ck, "trainer_state.json")))["global_step"]
    import math

    max_step = math.ceil(args.num_train_epochs * len(tr2.get_train_dataloader()) / args.gradient_accumulation_steps)      # main.py:599-604
    tr2.args.max_steps = resumed + max_step
    tr2.state.max_steps = max_step
    tr2.train(resume_from_checkpoint=ck)
    assert (mw.train.buf.data_ptr(), mw.frozen.buf.data_ptr()) == ptr_before, "resume must load into the live buffers"
    assert torch.equal(mw.train.buf, saved) and mw.frozen.buf.abs().sum() > 0
    assert e2.opt_step == 7 and torch.equal(e2.adam_v, torch.ones(n) * 0.5)
    assert tr2.state.global_step == resumed + max_step and e2.steps == max_step
    assert e2.lrs[0] == pytest.approx(1e-4) and e2.lrs[-1] < e2.lrs[0]        # the schedule restarts over the resumed span
    with pytest.raises(ValueError, match="another model configuration"):
        import dataclasses
        tr2.model_cfg = dataclasses.replace(cfg, hidden_size=cfg.hidden_size * 2)
        tr2._load_checkpoint(ck)
