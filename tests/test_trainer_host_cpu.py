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


# ------------------------------------------------------------------------------------------------------------------
# uneven dataset shards, per-group metric functions, global sample filter
# ------------------------------------------------------------------------------------------------------------------
def _uneven_worker(rank, world, port, q, n_rows, accum, epochs):
    import os

    import torch.distributed as dist

    from time_r1_b200 import rewards
    from time_r1_b200.trainer import GRPOArguments, TimeR1_Trainer

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rewards._iou_on_device = lambda mode, completions, solution, durations: [0.5] * len(completions)
    args = GRPOArguments(output_dir="/tmp/unused", num_generations=2, max_completion_length=4, beta=0.0, gradient_accumulation_steps=accum,
                         num_train_epochs=epochs, logging_steps=1000, save_strategy="no")
    tr = make_trainer(TimeR1_Trainer, args, [rewards.iou_timestamp_reward])
    tr.train_dataset = rows(n_rows)
    tr.world, tr.rank = world, rank

    class CollectiveEngine(StandInEngine):        # every optimizer step holds a collective, like Engine.optimizer_step at world > 1
        def optimizer_step(self, lr=None):
            t = torch.ones(1)
            dist.all_reduce(t)                   # would hang (-> test timeout) if the ranks disagreed on the step count
            return super().optimizer_step(lr)

    tr.engine = CollectiveEngine(2, 4)
    lens = [len(tr.get_train_dataloader(e)) for e in range(2)]
    seen = sorted(int(mb[0]["problem"].split()[1]) for mb in tr.get_train_dataloader(0))
    tr.train()
    q.put((rank, lens, tr.state.max_steps, tr.state.global_step, tr.engine.steps, seen))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rows,accum,epochs", [(7, 2, 3), (5, 1, 2), (9, 4, 2)])
def test_uneven_dataset_gives_every_rank_the_same_step_count_gloo(n_rows, accum, epochs):
    """len(dataset) % world != 0 (train_2k5.json: 2500 rows on 8 ranks with accum 2): the shuffled index list is padded by
    wrap-around like DistributedSampler, so loader lengths, max_steps and the number of optimizer steps (each of which
    holds a collective) agree on all ranks and train() returns on every rank."""
    import socket

    import torch.multiprocessing as mp

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_uneven_worker, args=(r, world, port, q, n_rows, accum, epochs)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in ps:
        p.join(60)
    (_, lens0, max0, gs0, st0, seen0), (_, lens1, max1, gs1, st1, seen1) = res
    L = -(-n_rows // world)
    assert lens0 == lens1 == [L, L]
    assert max0 == max1 == int(np.ceil(epochs * max(1, L // accum)))
    assert gs0 == gs1 == st0 == st1 == max0
    assert set(seen0) | set(seen1) == set(range(n_rows))          # every row is dealt out; the padding repeats rows, never drops one


def test_metric_funcs_are_computed_from_each_groups_own_completions(tmp_path, monkeypatch):
    """gradient_accumulation_steps=2: both prompt-groups of an optimizer step log metrics/<fn> from THEIR completions
    (timer1_trainer_ft.py:670-691), not from the last group's."""
    from time_r1_b200 import rewards
    from time_r1_b200.trainer import GRPOArguments, TimeR1_Trainer_ft

    monkeypatch.setattr(rewards, "_iou_on_device", lambda mode, completions, solution, durations: [0.5] * len(completions))
    calls = []

    def think_chars(completions, **kw):
        vals = [float(len(c)) for c in completions]
        calls.append((kw["problem"][0], vals))
        return vals

    args = GRPOArguments(output_dir=str(tmp_path), num_generations=2, max_completion_length=4, beta=0.0, gradient_accumulation_steps=2,
                         max_steps=1, logging_steps=1, save_strategy="no")
    tr = make_trainer(TimeR1_Trainer_ft, args, [rewards.iou_timestamp_reward], metric_funcs=[think_chars])

    class Tok(TemplateTokenizer):
        def batch_decode(self, ids, skip_special_tokens=True):
            return ["x" * (1 + int(r[0]) % 50) for r in np.asarray(ids)]

    tr.processing_class = Tok()
    tr.train()
    assert len(calls) == 2 and calls[0][0] != calls[1][0]           # one call per group, each with its own dataset row
    assert calls[0][1] != calls[1][1]
    want = (np.mean(calls[0][1]) + np.mean(calls[1][1])) / 2        # log() averages the per-group entries
    assert tr.state.log_history[0]["metrics/think_chars"] == pytest.approx(want)


def test_fix_vit_false_moves_the_vit_into_the_trainable_buffer():
    """fix_vit=False is the reference default (main.py:46-49; timer1_trainer.py:272-280 then leaves model.visual trainable): the
    ViT tensors join the trainable flat buffer (they get gradients, AdamW state and ZeRO shards like everything else), ahead of
    the merger so that the backward still finishes the buffer from its end to its start; checkpoints round-trip unchanged."""
    import inspect

    import torch

    from time_r1_b200 import config, trainer, weights

    src = inspect.getsource(trainer.TimeR1_Trainer.__init__)
    assert "fix_vit=fix_vit" in src and "train_vit=not fix_vit" in src and "NotImplementedError(\"fix_vit" not in src
    for fam in ("qwen2_5_vl", "qwen2_vl"):
        cfg = config.toy(fam)
        frozen = weights.ModelWeights(cfg, "cpu", world_size=2)
        train = weights.ModelWeights(cfg, "cpu", world_size=2, train_vit=True)
        assert train.frozen.numel == 0 and not train.frozen.specs
        vit_names = [s.name for s in frozen.frozen.specs]
        assert vit_names and all(n in train.train and n not in frozen.train for n in vit_names)
        lay = train.train.layout
        assert max(lay[n][0] for n in vit_names) < lay["visual.merger.ln_q.weight"][0] < lay["model.embed_tokens.weight"][0]
        assert train.train.numel % (weights.ALIGN * 2 * 4) == 0
        frozen.init_random(seed=3)
        train.load_state_dict(frozen.export_state_dict())
        a, b = frozen.export_state_dict(), train.export_state_dict()
        assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)


def test_global_batches_are_dealt_evenly_and_keep_their_content():
    """dist_utils.balanced_deal in the dataloader: every optimizer step still processes the SAME set of samples the plain strided
    deal would (exact-math: the DP-averaged gradient does not care which rank ran a sample), but per-rank token counts inside a
    step are even — the step ends when the slowest rank reaches the reduce-scatter."""
    from time_r1_b200.dist_utils import balanced_deal
    from time_r1_b200.trainer import GRPOArguments, TimeR1_Trainer

    world, accum, n = 4, 2, 32
    rng = np.random.default_rng(0)
    durs = rng.choice([9.0, 30.0, 147.0, 600.0], n)
    data = [dict(problem=f"event {i}", video_path=f"/v/{i}.mp4", solution=(1.0, 2.0), durations=float(d)) for i, d in enumerate(durs)]
    per_rank = {}
    for balance in (False, True):
        for r in range(world):
            args = GRPOArguments(num_generations=2, max_completion_length=4, beta=0.0, gradient_accumulation_steps=accum, fix_vit=True)
            args.balance_ranks = balance
            tr = make_trainer(TimeR1_Trainer, args, [])
            tr.train_dataset, tr.world, tr.rank = data, world, r
            per_rank[(balance, r)] = [int(mb[0]["problem"].split()[1]) for mb in tr.get_train_dataloader(0)]
    steps = len(per_rank[(True, 0)]) // accum
    crit = {False: 0.0, True: 0.0}      # sum over steps of the slowest rank's load = the epoch's critical path
    ideal = 0.0
    for s in range(steps):
        sets = {b: sorted(i for r in range(world) for i in per_rank[(b, r)][s * accum:(s + 1) * accum]) for b in (False, True)}
        assert sets[False] == sets[True], "balancing must not move a sample to another optimizer step"
        for b in (False, True):
            loads = [sum(TimeR1_Trainer._sample_cost(data[i]) for i in per_rank[(b, r)][s * accum:(s + 1) * accum]) for r in range(world)]
            crit[b] += max(loads)
        ideal += sum(loads) / world
    assert crit[True] < 0.93 * crit[False] and crit[True] < 1.2 * ideal, (crit, ideal)
    assert balanced_deal([5, 1, 1, 5], 2) == [0, 3, 1, 2] and sorted(balanced_deal(list(range(12)), 4)) == list(range(12))


@pytest.mark.parametrize("train_vit", [False, True])
@pytest.mark.parametrize("family", ["qwen2_5_vl", "qwen2_vl"])
def test_overlapped_reduce_scatter_never_takes_a_bucket_before_its_gradients_are_final(family, train_vit):
    """Engine._backward / _vit_backward call _reduce_ready(offset) as the backward walks the flat buffer from its end (lm_head)
    to its start (merger; with fix_vit=False the ViT blocks, then the patch-embed): replay that call sequence on the BucketPlan
    and check, for every bucket, that each tensor it overlaps was final when the bucket was handed to NCCL (the order in which the
    backward finishes tensors: lm_head, final norm, layers L-1..0, embedding scatter, merger, ViT blocks depth-1..0, patch-embed),
    that no bucket is handed out twice, and that optimizer_step's ready_buckets(0) collects the rest."""
    from time_r1_b200 import config, weights
    from time_r1_b200.dist_utils import BucketPlan

    cfg = config.toy(family)
    world = 2
    mw = weights.ModelWeights(cfg, "cpu", world_size=world, train_vit=train_vit)
    lay = mw.train.layout
    off = lambda n: lay[n][0]
    L, depth = cfg.num_hidden_layers, cfg.vit_depth
    # when is a tensor final?  (index into the sequence of _reduce_ready calls below; the call happens AFTER the work it names)
    calls = [("lm_head.weight", off("lm_head.weight")), ("model.norm.weight", off("model.norm.weight"))]
    calls += [(f"model.layers.{l}.", off(f"model.layers.{l}.input_layernorm.weight")) for l in range(L - 1, -1, -1)]
    if train_vit:
        calls.append(("visual.merger.", off("visual.merger.ln_q.weight")))        # after the embedding scatter and the merger backward
        calls += [(f"visual.blocks.{i}.", off(f"visual.blocks.{i}.norm1.weight")) for i in range(depth - 1, -1, -1)]
    calls.append(("<optimizer_step>", 0))

    def final_at(name):
        if name in ("model.embed_tokens.weight",) or name.startswith("visual.merger."):
            key = "visual.merger." if train_vit else "<optimizer_step>"
        elif name == "visual.patch_embed.proj.weight":
            key = "<optimizer_step>"
        elif name.startswith("visual.blocks."):
            key = ".".join(name.split(".")[:3]) + "."
        elif name.startswith("model.layers."):
            key = ".".join(name.split(".")[:3]) + "."
        else:
            key = name
        return next(i for i, (k, _) in enumerate(calls) if k == key)

    for n_buckets in (4, 7, 32):
        plan = BucketPlan(mw.train.numel, world, n_buckets=n_buckets)
        issued = {}
        done_prev = None
        for t, (_, done_from) in enumerate(calls):
            assert done_prev is None or done_from <= done_prev, "the backward walks the buffer from its end to its start"
            done_prev = done_from
            for b in plan.ready_buckets(done_from):
                assert b not in issued
                issued[b] = t
        assert sorted(issued) == list(range(len(plan)))
        for name, (o, shape) in lay.items():
            n = 1
            for d in shape:
                n *= d
            for b, (lo, hi) in enumerate(plan.bounds):
                if lo < o + n and hi > o:
                    assert issued[b] >= final_at(name), (name, b, n_buckets)
