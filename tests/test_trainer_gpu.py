"""Drop-in trainer surface on a toy model: reward protocol, metric keys, grad accumulation, callbacks, HF-format
checkpoint round trip (mirrors how main.py:573-625 drives the reference trainer)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class StubTokenizer:
    """ids -> text: id 100+k decodes to a template answer so rewards are non-degenerate; special ids are skipped."""
    eos_token_id = 2043

    def batch_decode(self, ids, skip_special_tokens=True):
        out = []
        for row in np.asarray(ids):
            toks = [int(t) for t in row if int(t) not in (2043, 2044)]
            s = (toks[0] % 17) * 1.0 if toks else 0.0
            e = s + 1.0 + (toks[1] % 13 if len(toks) > 1 else 0)
            tail = "" if (toks and toks[0] % 5 == 0) else "</answer>"
            out.append(f"<think>ids {toks[:3]}</think><answer>{s:.2f} to {e:.2f}{tail}")
        return out


def make_dataset(cfg, n=6):
    rng = np.random.default_rng(0)
    rows = []
    for i in range(n):
        grid = (2, 8, 12) if i % 2 == 0 else (2, 8, 8)
        N = grid[0] * grid[1] * grid[2]
        ids = np.array([5, 6, cfg.vision_start_token_id] + [cfg.video_token_id] * (N // 4) + [9] + list(rng.integers(10, 2000, 6)) + [7])
        rows.append(dict(problem=f"event {i}", video_path=f"/videos/{i}.mp4", solution=(2.0 + i, 9.0 + i), durations=30.0 + i,
                         task_type="tg", choices="", video_start=None, video_end=None, preprocessed_path="",
                         input_ids=ids, pixel_values_videos=rng.standard_normal((N, cfg.patch_dim)).astype(np.float32),
                         video_grid_thw=np.array([grid]), second_per_grid_ts=[1.0]))
    return rows


def test_trainer_runs_like_main_py(tmp_path):
    from src.time_r1 import TimeR1_Trainer
    from time_r1_b200 import config, rewards, weights
    from time_r1_b200.trainer import GRPOArguments, TrainerCallback, load_checkpoint

    cfg = config.toy()
    mw = weights.ModelWeights(cfg, "cuda")
    mw.init_random(seed=1, std=0.05)
    seen = {}

    def spy_reward(prompts, completions, **kw):
        seen["kw"] = sorted(kw)
        seen["n"] = (len(prompts), len(completions), len(kw["solution"]))
        return [float(len(c) % 3) for c in completions]

    class Stop(TrainerCallback):
        epochs = 0

        def on_epoch_end(self, args, state, control, **kw):
            Stop.epochs += 1
            if state.epoch >= 1:
                control.should_training_stop = True

    args = GRPOArguments(output_dir=str(tmp_path), num_generations=4, max_completion_length=6, temperature=1.0, beta=0.04,
                         learning_rate=1e-4, gradient_accumulation_steps=2, num_train_epochs=5, logging_steps=1, fix_vit=True,
                         save_strategy="epoch")
    tr = TimeR1_Trainer(model=(cfg, mw), reward_funcs=[rewards.iou_timestamp_reward_v2, rewards.format_reward, spy_reward],
                        metric_funcs=[], args=args, train_dataset=make_dataset(cfg), processing_class=StubTokenizer(),
                        callbacks=[Stop()])
    w0 = mw.train.buf.clone()
    tr.train()
    assert Stop.epochs == 1 and tr.state.global_step == 3            # 6 examples / accum 2, stopped after epoch 1
    assert seen["n"] == (4, 4, 4)
    for col in ("solution", "durations", "video_path", "problem", "task_type", "choices", "video_start", "video_end", "preprocessed_path"):
        assert col in seen["kw"], col
    log = tr.state.log_history[-1]
    for key in ("loss", "grad_norm", "learning_rate", "completion_length", "rewards/iou_timestamp_reward_v2",
                "rewards/format_reward", "rewards/spy_reward", "reward", "reward_std", "kl", "generation_entropy"):
        assert key in log, key
    assert 0 < log["completion_length"] <= 6 and log["kl"] >= -1e-6
    assert abs(tr.state.log_history[0]["learning_rate"] - 1e-4) < 1e-12 and tr.state.log_history[1]["learning_rate"] < 1e-4
    assert not torch.equal(w0, mw.train.buf), "weights must move"
    ck = os.path.join(str(tmp_path), "checkpoint-3")
    assert os.path.exists(os.path.join(ck, "config.json")) and os.path.exists(os.path.join(ck, "trainer_state.json"))
    cfg2, w2 = load_checkpoint(ck, "cuda")
    assert cfg2.hidden_size == cfg.hidden_size and cfg2.vit_intermediate == cfg.vit_intermediate
    assert torch.equal(w2.train.buf, mw.train.buf) and torch.equal(w2.frozen.buf, mw.frozen.buf)
    tr.save_model(str(tmp_path / "final"))
    hf = json.load(open(tmp_path / "final" / "config.json"))
    assert hf["architectures"] == ["Qwen2_5_VLForConditionalGeneration"]


def test_trainer_with_fix_vit_false_trains_and_round_trips_the_tower(tmp_path):
    """--fix_vit false (main.py:46-49 default): the trainer builds / accepts train_vit weights, the ViT tensors move under AdamW,
    the reference copy keeps the initial tower, and the checkpoint reloads (into either layout) bit-equal."""
    from src.time_r1 import TimeR1_Trainer
    from time_r1_b200 import config, rewards, weights
    from time_r1_b200.trainer import GRPOArguments, load_checkpoint

    cfg = config.toy()
    mw = weights.ModelWeights(cfg, "cuda", train_vit=True)
    mw.init_random(seed=1, std=0.05)
    args = GRPOArguments(output_dir=str(tmp_path), num_generations=4, max_completion_length=6, temperature=1.0, beta=0.04,
                         learning_rate=1e-3, gradient_accumulation_steps=2, max_steps=2, logging_steps=1, fix_vit=False, save_strategy="no")
    with pytest.raises(ValueError, match="train_vit"):
        TimeR1_Trainer(model=(cfg, weights.ModelWeights(cfg, "cuda")), reward_funcs=[rewards.format_reward], args=args,
                       train_dataset=make_dataset(cfg), processing_class=StubTokenizer())
    tr = TimeR1_Trainer(model=(cfg, mw), reward_funcs=[rewards.iou_timestamp_reward_v2, rewards.format_reward], metric_funcs=[],
                        args=args, train_dataset=make_dataset(cfg), processing_class=StubTokenizer())
    name = "visual.blocks.1.attn.qkv.weight"
    w0 = mw.train[name].clone()
    tr.train()
    assert tr.state.global_step == 2 and np.isfinite(tr.state.log_history[-1]["loss"]) and tr.state.log_history[-1]["grad_norm"] > 0
    assert not torch.equal(mw.train[name], w0) and torch.equal(tr.engine.ref[name], w0)
    tr.save_model(str(tmp_path / "final"))
    for train_vit in (True, False):
        _, w2 = load_checkpoint(str(tmp_path / "final"), "cuda", train_vit=train_vit)
        a, b = mw.export_state_dict(), w2.export_state_dict()
        assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)


def test_trainer_rejects_unsupported_paths():
    from time_r1_b200 import config, rewards, weights
    from time_r1_b200.trainer import GRPOArguments, TimeR1_Trainer

    cfg = config.toy()
    mw = weights.ModelWeights(cfg, "cuda")
    mw.init_random(seed=1)
    tr = TimeR1_Trainer(model=(cfg, mw), reward_funcs=rewards.format_reward, args=GRPOArguments(num_generations=2, max_completion_length=4, beta=0.0, fix_vit=True),
                        train_dataset=[], processing_class=StubTokenizer())
    with pytest.raises(ValueError):
        tr.compute_loss(tr.model, [{}, {}])
    with pytest.raises(NotImplementedError):
        tr.compute_loss(tr.model, [dict(problem="x", video_path="/nope.mp4")])
    with pytest.raises(NotImplementedError):
        TimeR1_Trainer(model=(cfg, mw), reward_funcs=["some/reward-model"], args=GRPOArguments(fix_vit=True))


def test_ft_trainer_on_the_real_engine_with_finetune_py_rows(tmp_path):
    """TimeR1_Trainer_ft end to end on the CUDA engine from finetune.py-layout rows (video_inputs = [[frames]], video_kwargs =
    [{"fps": [f]}], timer1_trainer_ft.py:551-566): host tokenise + pad expansion, device patchify, rollout, scoring, metric
    functions per group, clip-ratio metrics, optimizer step."""
    from src.time_r1 import TimeR1_Trainer_ft
    from test_dropin_rows_cpu import FakeProcessor
    from time_r1_b200 import config, rewards, weights
    from time_r1_b200.trainer import GRPOArguments

    cfg = config.toy()
    mw = weights.ModelWeights(cfg, "cuda")
    mw.init_random(seed=1, std=0.05)

    class Proc(FakeProcessor, StubTokenizer):
        pass

    g = torch.Generator().manual_seed(0)
    data = [dict(problem=f"event {i}", video_path=f"/v/{i}.mp4", video_start=None, video_end=None, solution=(2.0 + i, 9.0 + i), durations=30.0,
                 preprocessed_path=f"/pre/{i}", use_preprocessed=[True],
                 video_inputs=[[torch.randint(0, 256, (4, 3, 112, 112 + 56 * (i % 2)), generator=g).float()]], video_kwargs=[{"fps": [2.0]}])
            for i in range(4)]
    args = GRPOArguments(output_dir=str(tmp_path), num_generations=4, max_completion_length=6, temperature=1.0, beta=0.04, learning_rate=1e-4,
                         gradient_accumulation_steps=2, max_steps=2, logging_steps=1, fix_vit=True, save_strategy="no", prompt_type="v2")
    tr = TimeR1_Trainer_ft(model=(cfg, mw), reward_funcs=[rewards.iou_timestamp_reward, rewards.format_reward],
                           metric_funcs=list(rewards.metric_funcs_registry.values()), args=args, train_dataset=data, processing_class=Proc(cfg))
    w0 = mw.train.buf.clone()
    tr.train()
    assert tr.state.global_step == 2 and not torch.equal(w0, mw.train.buf)
    log = tr.state.log_history[-1]
    for key in ["loss", "grad_norm", "completion_length", "rewards/iou_timestamp_reward", "rewards/format_reward", "reward", "reward_std", "kl",
                "generation_entropy", "clip_ratio/low_mean", "clip_ratio/low_min", "clip_ratio/high_mean", "clip_ratio/high_max",
                "clip_ratio/region_mean"] + [f"metrics/{k}" for k in rewards.metric_funcs_registry]:
        assert key in log and np.isfinite(log[key]), key
    assert "provide the start and end times" in tr.processing_class.seen[0]["content"][0]["text"].lower()      # prompt v2
