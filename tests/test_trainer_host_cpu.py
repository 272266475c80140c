"""Host logic of the drop-in trainers on the CPU: the train loop, reward / metric protocol, metric key names, linear LR
decay, grad accumulation, callbacks and the stop flag are driven against a stand-in engine that returns CPU tensors of the
engine's output contract (Engine.grpo_multi_step / optimizer_step).  The numerics live in tests/test_trainer_gpu.py; this
file keeps the Python that surrounds them honest where no GPU is available (timer1_trainer.py:512-793, main.py:497-625)."""
import collections

import numpy as np
import pytest
import torch


class StandInEngine:
    """Output contract of Engine.grpo_multi_step: per prompt-group a dict with loss [1], stats [8] (index 1 = kl, 2 =
    entropy, 3 = completion length, 5 / 6 = low / high clip fraction), rewards_per_func [G, n], rewards [G], reward_std."""

    def __init__(self, G, C):
        self.G, self.C, self.steps, self.lrs, self.seen_prompts = G, C, 0, [], []

    def grpo_multi_step(self, prompts, reward_fns, seed=None, loss_scale=1.0, vit_prefetch=None, **kw):
        outs = []
        for p, rf in zip(prompts, reward_fns):
            self.seen_prompts.append(len(p.input_ids))
            rng = np.random.default_rng(seed + len(outs))
            comp = rng.integers(100, 2000, (self.G, self.C)).astype(np.int64)
            assert getattr(rf, "device", None) is None, "no CUDA here: the host protocol must have been chosen"
            rpf = torch.from_numpy(np.asarray(rf(comp), np.float32).reshape(self.G, -1))
            rewards = rpf.sum(1)
            stats = torch.tensor([0.0, 0.01, 7.5, float(self.C), 0.0, 0.125, 0.25, 0.0])
            outs.append(dict(loss=torch.tensor([0.5 * loss_scale]), stats=stats, rewards_per_func=rpf, rewards=rewards,
                             reward_std=rewards.std().reshape(1)))
        return outs

    def prefetch_vit(self, prompts):
        return ("handle", len(prompts))

    def optimizer_step(self, lr=None):
        self.steps += 1
        self.lrs.append(lr)
        return torch.tensor(1.25)


class TemplateTokenizer:
    eos_token_id = 2043

    def batch_decode(self, ids, skip_special_tokens=True):
        out = []
        for row in np.asarray(ids):
            s = int(row[0]) % 17
            out.append(f"<think>step one, because <timestep>{s} to {s + 2}</timestep>\nnotice</think><answer>{s}.50 to {s + 3 + int(row[1]) % 5}.00</answer>")
        return out


def rows(n=6):
    rng = np.random.default_rng(0)
    return [dict(problem=f"event {i}", video_path=f"/v/{i}.mp4", solution=(2.0 + i, 9.0 + i), durations=30.0, task_type="tg", choices="",
                 video_start=None, video_end=None, preprocessed_path="", input_ids=rng.integers(0, 100, 20 + i),
                 pixel_values_videos=np.zeros((8, 4), np.float32), video_grid_thw=np.array([[2, 2, 2]]), second_per_grid_ts=[1.0])
            for i in range(n)]


def make_trainer(cls, args, reward_funcs, metric_funcs=(), callbacks=()):
    from time_r1_b200.trainer import TrainerControl, TrainerState

    tr = cls.__new__(cls)              # everything __init__ sets up that does not need CUDA
    tr.args, tr.reward_funcs, tr.metric_funcs = args, list(reward_funcs), list(metric_funcs)
    tr.train_dataset, tr.processing_class, tr.vision_processor, tr.callbacks = rows(), TemplateTokenizer(), None, list(callbacks)
    tr.num_generations, tr.max_completion_length, tr.beta = args.num_generations, args.max_completion_length, args.beta
    tr.prompt_type, tr.use_grpo = args.prompt_type, args.use_grpo
    tr._metrics, tr.state, tr.control = collections.defaultdict(list), TrainerState(), TrainerControl()
    tr.world, tr.rank, tr.device = 1, 0, torch.device("cpu")
    tr.state.is_world_process_zero = True
    tr.run = type("Run", (), dict(seed=42, fix_vit=True))()
    tr.engine = StandInEngine(args.num_generations, args.max_completion_length)
    return tr


def test_train_loop_protocol_and_metric_keys(tmp_path, monkeypatch):
    from time_r1_b200 import rewards
    from time_r1_b200.trainer import GRPOArguments, TimeR1_Trainer_ft, TrainerCallback

    # the IoU arithmetic is a CUDA kernel in the product: here only the protocol around it is under test
    monkeypatch.setattr(rewards, "_iou_on_device", lambda mode, completions, solution, durations: [0.5] * len(completions))
    seen = {}

    def spy_reward(prompts, completions, **kw):
        seen["kw"] = sorted(kw)
        seen["prompt_text"] = prompts[0][0]["content"][0]["text"]
        return [1.0] * len(completions)

    class Stop(TrainerCallback):
        def on_epoch_end(self, args, state, control, **kw):
            seen.setdefault("epochs", []).append(state.epoch)
            assert kw["trainer"] is not None
            if state.epoch >= 2:
                control.should_training_stop = True

    args = GRPOArguments(output_dir=str(tmp_path), num_generations=4, max_completion_length=6, beta=0.04, learning_rate=1e-4,
                         gradient_accumulation_steps=2, num_train_epochs=5, logging_steps=1, save_strategy="no", prompt_type="v3")
    tr = make_trainer(TimeR1_Trainer_ft, args, [rewards.iou_timestamp_reward_v2, rewards.format_reward, spy_reward],
                      metric_funcs=list(rewards.metric_funcs_registry.values()), callbacks=[Stop()])
    tr.train()
    assert seen["epochs"] == [1.0, 2.0] and tr.state.global_step == 6 and tr.engine.steps == 6      # 6 rows / accum 2, two epochs
    assert tr.state.max_steps == 15                                                                 # ceil(5 epochs * 6 / 2)
    assert tr.engine.seen_prompts and set(tr.engine.seen_prompts) <= set(range(20, 26))
    assert "Step-by-step Analysis" in seen["prompt_text"] and "event" in seen["prompt_text"]        # prompt v3
    for col in ("solution", "durations", "video_path", "problem", "task_type", "choices", "video_start", "video_end", "preprocessed_path"):
        assert col in seen["kw"], col
    for banned in ("input_ids", "pixel_values_videos", "video_grid_thw"):
        assert banned not in seen["kw"]
    log = tr.state.log_history[0]
    want = {"loss", "grad_norm", "learning_rate", "epoch", "step", "completion_length", "rewards/iou_timestamp_reward_v2", "rewards/format_reward",
            "rewards/spy_reward", "reward", "reward_std", "kl", "generation_entropy", "clip_ratio/low_mean", "clip_ratio/low_min",
            "clip_ratio/high_mean", "clip_ratio/high_max", "clip_ratio/region_mean"} | {f"metrics/{k}" for k in rewards.metric_funcs_registry}
    assert set(log) == want, set(log) ^ want
    assert log["loss"] == 0.5 and log["rewards/format_reward"] == 1.0 and log["reward"] == 2.5 and log["completion_length"] == 6.0
    assert log["metrics/reward_timestep_pair"] == pytest.approx(0.2) and log["metrics/reward_paragraph_structure"] == pytest.approx(0.1)
    assert log["clip_ratio/region_mean"] == 0.375 and log["kl"] == pytest.approx(0.01)
    lrs = [l["learning_rate"] for l in tr.state.log_history]
    assert lrs[0] == 1e-4 and all(a > b for a, b in zip(lrs, lrs[1:])) and lrs == tr.engine.lrs   # linear decay, no warm-up


def test_base_trainer_rejects_other_prompt_types_and_keeps_metric_funcs_silent(tmp_path, monkeypatch):
    from time_r1_b200 import rewards
    from time_r1_b200.trainer import GRPOArguments, TimeR1_Trainer

    monkeypatch.setattr(rewards, "_iou_on_device", lambda mode, completions, solution, durations: [0.25] * len(completions))
    args = GRPOArguments(output_dir=str(tmp_path), num_generations=2, max_completion_length=4, beta=0.0, gradient_accumulation_steps=1,
                         max_steps=2, logging_steps=1, save_strategy="no")
    tr = make_trainer(TimeR1_Trainer, args, [rewards.iou_timestamp_reward], metric_funcs=list(rewards.metric_funcs_registry.values()))
    tr.train()
    log = tr.state.log_history[-1]
    assert tr.state.global_step == 2 and "kl" not in log and not any(k.startswith(("metrics/", "clip_ratio/")) for k in log)
    assert log["rewards/iou_timestamp_reward"] == 0.25
    tr.prompt_type = "v2"
    with pytest.raises(ValueError):
        tr.compute_loss(None, [tr.train_dataset[0]])
    with pytest.raises(ValueError):
        tr.compute_loss(None, [tr.train_dataset[0], tr.train_dataset[1]])          # per-device batch must be 1
