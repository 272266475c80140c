"""Drop-in `TimeR1_Trainer` — same constructor, reward-function protocol, metric names and train()/save_model()/log()
surface as src/time_r1/rl/timer1_trainer.py:100-857, driven by the B200 engine instead of HF Trainer + DeepSpeed.

What is mirrored (reference file:line):
  * ctor kwargs                                   timer1_trainer.py:184-206
  * ε_low = ε_high = 0.2 regardless of args       :388-393
  * one prompt-group per micro-step (bs forced 1) :548-551
  * reward protocol f(prompts=, completions=, **dataset columns repeated G times) -> list[float]   :685-698
  * metric keys completion_length, rewards/<fn>, reward, reward_std, kl, generation_entropy  :740-777, log() :784-793
  * loss / grad-accum scaling, clip 1.0, AdamW, linear LR decay    HF trainer.py:1927,2493,1760
  * callbacks: on_train_begin / on_step_end / on_epoch_end(control.should_training_stop) / on_log / on_train_end
    (main.py:497-539), `.state.max_steps`, `.args`, `.get_train_dataloader()` (main.py:599-618)
What differs by design: the model never lives in torch modules — `model` is a checkpoint directory (HF safetensors,
released-checkpoint names), a preset name ("random:qwen2.5-vl-7b", "random:toy") or a (ModelConfig, ModelWeights) pair.
Video decoding / HF processor (timer1_trainer.py:518-556) is an injection point (`vision_processor`); dataset rows that
already carry `input_ids` + `pixel_values_videos` + `video_grid_thw` (the `_ft` variant's pre-decoded layout) skip it.
"""
from __future__ import annotations

import json
import math
import os
import random
from collections import defaultdict
from dataclasses import asdict, dataclass, field
from typing import Any, Callable, Optional, Sequence, Union

import numpy as np

from .config import PRESETS, ModelConfig, RunConfig, from_hf_config

SYSTEM_PROMPT = "You are a video analysis expert."
QUESTION_TEMPLATE_TG_v1 = (
    'To accurately pinpoint the event "[EVENT]" in the video, determine the precise time period of the event.\n\n'
    "Output your thought process within the <think> </think> tags, including analysis with either specific time "
    "ranges (xx.xx to xx.xx) in <timestep> </timestep> tags.\n\n"
    'Then, provide the start and end times (in seconds, precise to two decimal places) in the format "start time to '
    'end time" within the <answer> </answer> tags. For example: "12.54 to 17.83".'
)
# timer1_trainer_ft.py:67-71 (prompt_type v2: answer only; v3: step-by-step <timestep> analysis)
QUESTION_TEMPLATE_TG_v2 = (
    'To accurately pinpoint the event "[EVENT]" in the video, determine the precise time period of the event.\n'
    '\n'
    'Provide the start and end times (in seconds, precise to two decimal places) in the format "start time to end time" within the <answer> </answer> tags. For example: "12.54 to 17.83".'
)
QUESTION_TEMPLATE_TG_v3 = (
    'Carefully analyze the video content to determine the precise time period during which "[EVENT]" occurs.  Within the `<think>` tags, provide a detailed description of your thought process, following the format below:\n'
    '```\n'
    '<think>\n'
    'Step-by-step Analysis:\n'
    '<timestep>Time period 1 (start time to end time)</timestep>: Describe the video content within this time period and determine if it is related to "[EVENT]".\n'
    '<timestep>Time period 2 (start time to end time)</timestep>: Describe the video content within this time period and determine if it is related to "[EVENT]".\n'
    'Based on the above analysis, state the precise time period during which "[EVENT]" occurs.\n'
    '</think>\n'
    '```\n'
    'Finally, in the `<answer>` tags, provide the start and end times of "[EVENT]" in the format "start time to end time" (in seconds, precise to two decimal places). For example: "12.54 to 17.83".\n'
    '```\n'
    '<answer>\n'
    'start time to end time\n'
    '</answer>\n'
    '```'
)
QUESTION_TEMPLATES = {"v1": QUESTION_TEMPLATE_TG_v1, "v2": QUESTION_TEMPLATE_TG_v2, "v3": QUESTION_TEMPLATE_TG_v3}


@dataclass
class GRPOArguments:
    """Union of the TrainingArguments / GRPOConfig / MY_GRPOConfig fields the reference consumes (SURVEY §5 config row;
    flags of scripts/posttrain/train_rl.sh:18-45).  Any object with these attributes (e.g. a real trl GRPOConfig) works."""
    output_dir: str = "./logs/timer1"
    num_generations: int = 8
    max_prompt_length: Optional[int] = 8192          # stored, never enforced (timer1_trainer.py:364)
    max_completion_length: int = 200
    temperature: float = 0.9                          # trl 0.17 default; scripts pass 1.0
    beta: float = 0.04
    learning_rate: float = 1e-6
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_epsilon: float = 1e-8
    weight_decay: float = 0.0
    max_grad_norm: float = 1.0
    per_device_train_batch_size: int = 1
    gradient_accumulation_steps: int = 1
    num_train_epochs: float = 1.0
    max_steps: int = -1
    logging_steps: int = 1
    save_steps: int = 500
    save_strategy: str = "steps"
    save_only_model: bool = False
    seed: int = 42
    data_seed: Optional[int] = None
    bf16: bool = True
    gradient_checkpointing: bool = True
    fix_vit: bool = False
    slide_window: bool = False
    max_window_layers: int = 2
    sliding_window_length: int = 4096
    prompt_type: str = "v1"
    use_grpo: bool = False
    top_k: int = 50                                   # HF 4.51.1 GenerationConfig default
    resume_from_checkpoint: Optional[str] = None
    push_to_hub: bool = False
    report_to: str = "none"
    run_name: Optional[str] = None
    logging_dir: Optional[str] = None
    deepspeed: Optional[str] = None                   # accepted and ignored: sharding is built in
    device_rewards: bool = True                       # registry rewards parsed on the device (rewards.DeviceRewards)
    model_init_kwargs: Optional[dict] = None


class TrainerState:
    def __init__(self):
        self.global_step = 0
        self.epoch = 0.0
        self.max_steps = 0
        self.is_world_process_zero = True
        self.log_history: list = []


class TrainerControl:
    def __init__(self):
        self.should_training_stop = False
        self.should_save = False
        self.should_log = False


class TrainerCallback:
    """Same hook names as transformers.TrainerCallback; every hook is optional."""

    def on_train_begin(self, args, state, control, **kw): ...
    def on_step_end(self, args, state, control, **kw): ...
    def on_epoch_end(self, args, state, control, **kw): ...
    def on_log(self, args, state, control, **kw): ...
    def on_train_end(self, args, state, control, **kw): ...


def _is_conversational(example: dict) -> bool:
    p = example.get("prompt")
    return isinstance(p, list) and len(p) > 0 and isinstance(p[0], dict) and "role" in p[0]


class TimeR1_Trainer:
    _logs_metric_funcs = False       # the base trainer stores metric_funcs and never calls them (timer1_trainer.py:330)

    def __init__(
        self,
        model: Union[str, tuple],
        reward_funcs: Union[Callable, Sequence[Callable]],
        metric_funcs: Union[Callable, Sequence[Callable], None] = None,
        args: Any = None,
        train_dataset: Optional[Sequence[dict]] = None,
        eval_dataset: Optional[Any] = None,
        processing_class: Optional[Any] = None,
        reward_processing_classes: Optional[Any] = None,
        callbacks: Optional[list] = None,
        optimizers: tuple = (None, None),
        peft_config: Optional[Any] = None,
        max_pixels: Optional[int] = 12845056,
        min_pixels: Optional[int] = 3136,
        attn_implementation: str = "flash_attention_2",
        vision_processor: Optional[Callable[[dict], dict]] = None,
    ):
        import torch

        from . import engine as _engine, lib, weights as _weights

        if peft_config is not None:
            raise NotImplementedError("LoRA/peft is never enabled by the reference scripts and is not built (SURVEY §2.2 T8)")
        if optimizers != (None, None):
            raise NotImplementedError("custom optimizers are not supported: the sharded AdamW kernel is built in")
        self.args = args if args is not None else GRPOArguments()
        a = self.args
        if not isinstance(reward_funcs, (list, tuple)):
            reward_funcs = [reward_funcs]
        for rf in reward_funcs:
            if isinstance(rf, str):
                raise NotImplementedError("reward models given as model ids are unused by the reference scripts (timer1_trainer.py:324-328)")
        self.reward_funcs = list(reward_funcs)
        self.metric_funcs = list(metric_funcs) if metric_funcs else []
        self.train_dataset, self.eval_dataset = train_dataset, eval_dataset
        self.processing_class = processing_class
        self.vision_processor = vision_processor
        self.callbacks = list(callbacks or [])
        self.max_pixels, self.min_pixels = max_pixels, min_pixels
        self.num_generations = a.num_generations
        self.max_completion_length = a.max_completion_length
        self.max_prompt_length = getattr(a, "max_prompt_length", None)
        self.beta = a.beta
        self.use_grpo = getattr(a, "use_grpo", False)
        self.prompt_type = getattr(a, "prompt_type", "v1")
        self.epsilon_low, self.epsilon_high = 0.2, 0.2       # hard-coded in the reference (:388-393)
        self._metrics: dict = defaultdict(list)
        self.state, self.control = TrainerState(), TrainerControl()

        # ---- distributed plumbing (one process per GPU, NCCL) ----
        import torch.distributed as dist

        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("TimeR1_Trainer (time_r1_b200) needs a CUDA device: there is no CPU fallback")
        torch.cuda.set_device(local)
        self.device = torch.device("cuda", local)
        pg = None
        if self.world > 1:
            if not dist.is_initialized():
                dist.init_process_group("nccl", device_id=self.device)
            pg = dist.group.WORLD
        self.state.is_world_process_zero = self.rank == 0
        lib.load()

        # ---- model ----
        # fix_vit (main.py:46-49, default False; every shipped script passes true): false leaves model.visual trainable
        # (timer1_trainer.py:272-280) -> the ViT patch-embed and blocks join the trainable flat buffer and receive gradients
        fix_vit = bool(getattr(a, "fix_vit", False))
        if isinstance(model, tuple):
            self.model_cfg, self.weights = model
            if bool(getattr(self.weights, "train_vit", False)) != (not fix_vit):
                raise ValueError(f"fix_vit={fix_vit} needs a (ModelConfig, ModelWeights) pair built with train_vit={not fix_vit}")
        elif isinstance(model, str) and model.startswith("random:"):
            self.model_cfg = PRESETS[model.split(":", 1)[1]]()
            self.weights = _weights.ModelWeights(self.model_cfg, self.device, world_size=self.world, train_vit=not fix_vit)
            self.weights.init_random(seed=getattr(a, "seed", 42))
        elif isinstance(model, str):
            self.model_cfg, self.weights = load_checkpoint(model, self.device, self.world, train_vit=not fix_vit)
            if self.processing_class is None:
                try:
                    from transformers import AutoProcessor

                    self.processing_class = AutoProcessor.from_pretrained(model)
                except Exception:
                    self.processing_class = None
        else:
            raise TypeError("model must be a checkpoint path, 'random:<preset>' or a (ModelConfig, ModelWeights) pair")
        self.run = RunConfig(
            num_generations=a.num_generations, max_completion_length=a.max_completion_length,
            temperature=a.temperature, top_k=getattr(a, "top_k", 50), beta=a.beta, epsilon_low=0.2, epsilon_high=0.2,
            use_grpo=self.use_grpo, learning_rate=a.learning_rate, adam_beta1=getattr(a, "adam_beta1", 0.9),
            adam_beta2=getattr(a, "adam_beta2", 0.999), adam_eps=getattr(a, "adam_epsilon", 1e-8),
            weight_decay=getattr(a, "weight_decay", 0.0), max_grad_norm=getattr(a, "max_grad_norm", 1.0),
            gradient_accumulation_steps=a.gradient_accumulation_steps, fix_vit=fix_vit, seed=getattr(a, "seed", 42))
        if getattr(a, "deepspeed", None) and self.rank == 0:
            print(f"[time_r1_b200] note: --deepspeed {a.deepspeed} is accepted for CLI compatibility and NOT applied: parameters and "
                  "gradients are replicated bf16, fp32 master weights + AdamW state are sharded 1/world per rank in HBM (ZeRO-2 "
                  "semantics, no CPU offload)")
        self.engine = _engine.Engine(self.model_cfg, self.weights, self.run, with_ref=a.beta != 0.0, rank=self.rank,
                                     world_size=self.world, process_group=pg)
        self.model = self.engine
        self.eos_token_id = getattr(getattr(self.processing_class, "tokenizer", self.processing_class), "eos_token_id", None) \
            or self.model_cfg.eos_token_id

    # ------------------------------------------------------------------ data
    def _epoch_indices(self, epoch: int = 0) -> list:
        """This rank's dataset indices for one epoch.  The shuffled list is padded by wrap-around to a multiple of the world
        size before it is dealt out (what torch's DistributedSampler / accelerate do under HF Trainer), so EVERY rank sees
        the same number of micro-batches: step counts, max_steps and the collectives of optimizer_step stay matched."""
        n = len(self.train_dataset)
        idx = list(range(n))
        seed = getattr(self.args, "data_seed", None)
        seed = getattr(self.args, "seed", 42) if seed is None else seed
        random.Random(seed + epoch).shuffle(idx)
        total = -(-n // self.world) * self.world
        idx = (idx * (total // max(1, n) + 1))[:total] if n else idx
        if self.world > 1 and getattr(self.args, "balance_ranks", True):
            # inside every global batch (world x gradient_accumulation_steps samples, the SAME set the plain strided deal would
            # process in that optimizer step) give each rank an even share of the work: the step ends when the slowest rank
            # reaches the gradient reduce-scatter, and TimeRFT clips differ 2-3x in tokens.  The DP-averaged gradient does not
            # depend on which rank ran which sample.
            from .dist_utils import balanced_deal

            chunk = self.world * max(1, getattr(self.args, "gradient_accumulation_steps", 1))
            out = []
            for a in range(0, total, chunk):
                part = idx[a:a + chunk]
                order = balanced_deal([self._sample_cost(self.train_dataset[i]) for i in part], self.world)
                out.extend(part[j] for j in order)
            idx = out
        return idx[self.rank:total:self.world]

    @staticmethod
    def _sample_cost(row: dict) -> float:
        """Relative cost of one prompt-group ~ its prompt tokens: known exactly for pre-processed rows, else from the clip
        duration the way fetch_video_v3 samples it (2 fps, 4..768 frames, vision_process.py:154-199; the per-frame pixel
        budget shrinks as frames grow, so tokens saturate around 3584 x 28 x 28 / 784 px)."""
        if row.get("input_ids") is not None:
            return float(len(row["input_ids"]))
        if row.get("video_grid_thw") is not None:
            g = np.asarray(row["video_grid_thw"]).reshape(-1, 3)
            return float((g[:, 0] * g[:, 1] * g[:, 2]).sum() // 4)
        dur = row.get("durations", row.get("duration"))
        try:
            frames = min(768.0, max(4.0, 2.0 * float(dur)))
        except (TypeError, ValueError):
            return 1.0
        return min(frames * 144.0, 3584.0 + frames * 7.5) + 120.0

    def get_train_dataloader(self, epoch: int = 0) -> list:
        """This rank's micro-batches for one epoch: lists of `per_device_train_batch_size` (=1) examples; the same length
        on every rank (main.py:599-605 derives max_steps from it)."""
        return [[self.train_dataset[i]] for i in self._epoch_indices(epoch)]

    _prompt_types = ("v1",)          # timer1_trainer.py:490-492 only knows v1

    def make_conversation_video(self, example: dict) -> list:
        if self.prompt_type not in self._prompt_types:
            raise ValueError(f"prompt_type {self.prompt_type!r} is not one of {self._prompt_types} for {type(self).__name__}")
        text = QUESTION_TEMPLATES[self.prompt_type].replace("[EVENT]", example["problem"])
        return [{"role": "user", "content": [
            {"type": "text", "text": text},
            {"type": "video", "video": example["video_path"], "video_start": example.get("video_start"),
             "video_end": example.get("video_end"), "total_pixels": 3584 * 28 * 28, "min_pixels": 16 * 28 * 28}]}]

    def _host_prompt_inputs(self, example: dict, frames, fps: float) -> dict:
        """What the reference does between the video reader and `generate` (timer1_trainer.py:536-556,
        timer1_trainer_ft.py:543-566), minus the pixel work: chat template -> tokenise -> expand the single <|video_pad|>
        to one token per merged patch (what the HF processor does from video_grid_thw).  The frames [T, 3, H, W] (0..255,
        already resized by fetch_video_v3 to multiples of 28 inside the pixel budget, vision_process.py:440-475) stay
        pixels: rescale / normalise / patchify run on the device (t1_video_patchify).  second_per_grid_ts =
        temporal_patch_size / fps, as the HF processor derives it from `fps=`."""
        import torch

        cfg = self.model_cfg
        pc = self.processing_class
        if pc is None or not hasattr(pc, "apply_chat_template"):
            raise RuntimeError("a processing_class with apply_chat_template() and a tokenizer is required to build prompts from raw rows")
        fr = frames if isinstance(frames, torch.Tensor) else torch.as_tensor(np.asarray(frames))
        if fr.dim() != 4 or fr.shape[1] != cfg.in_channels:
            raise ValueError(f"video frames must be [T, {cfg.in_channels}, H, W], got {tuple(fr.shape)}")
        T, _, H, W = fr.shape
        unit = cfg.patch_size * cfg.spatial_merge_size
        if H % unit or W % unit:
            raise ValueError(f"frame size {H}x{W} is not a multiple of {unit}: resize with smart_resize first (vision_process.py:60-90)")
        if T % cfg.temporal_patch_size:                       # the HF video processor repeats the last frame
            fr = torch.cat([fr, fr[-1:].expand(cfg.temporal_patch_size - T % cfg.temporal_patch_size, -1, -1, -1)], 0)
            T = fr.shape[0]
        grid = (T // cfg.temporal_patch_size, H // cfg.patch_size, W // cfg.patch_size)
        nv = grid[0] * grid[1] * grid[2] // cfg.merge_unit
        text = pc.apply_chat_template(self.make_conversation_video(example), tokenize=False, add_generation_prompt=True)
        tok = getattr(pc, "tokenizer", pc)
        ids = tok(text, add_special_tokens=False)["input_ids"]
        ids = np.asarray(ids[0] if len(ids) and isinstance(ids[0], (list, tuple, np.ndarray)) else ids, dtype=np.int64).reshape(-1)
        where = np.nonzero(ids == cfg.video_token_id)[0]
        if len(where) != 1:
            raise ValueError(f"the chat template must produce exactly one video placeholder token ({cfg.video_token_id}), found {len(where)}")
        ids = np.concatenate([ids[:where[0]], np.full(nv, cfg.video_token_id, np.int64), ids[where[0] + 1:]])
        if fr.dtype != torch.uint8:
            fr = fr.to(torch.float32)
        return dict(input_ids=ids, video_frames=fr, video_grid_thw=np.array([grid]),
                    second_per_grid_ts=[cfg.temporal_patch_size / float(fps)])

    def _default_vision_processor(self, example: dict) -> dict:
        """main.py rows (`video_path` only, main.py:455-468): the reference's own reader, `process_vision_info_v3`
        (src/utils/vision_process.py:547-578, decord + bicubic resize on the host) when that package is importable — this
        repository's `src/` is a namespace package, so with the reference checkout on PYTHONPATH `src.utils` resolves there."""
        try:
            from src.utils import process_vision_info_v3
        except Exception as e:
            raise NotImplementedError(
                "raw video decoding (timer1_trainer.py:518-534) is the reference's host-side reader: put the reference checkout "
                "on PYTHONPATH (src.utils.process_vision_info_v3 needs decord / torchvision), or pass vision_processor="
                "callable(example) -> {input_ids, pixel_values_videos | video_frames, video_grid_thw, second_per_grid_ts}, or "
                f"pre-processed dataset rows ({type(e).__name__}: {e})") from e
        messages = [{"role": "user", "content": [{"type": "video", "video": example["video_path"], "total_pixels": 3584 * 28 * 28,
                                                  "min_pixels": 16 * 28 * 28}]}]                      # timer1_trainer.py:518-530
        _, video_inputs, video_kwargs = process_vision_info_v3([messages], return_video_kwargs=True)
        return self._host_prompt_inputs(example, video_inputs[0], video_kwargs["fps"][0])

    @staticmethod
    def _unwrap_ft(example: dict):
        """`_ft` rows (finetune.py:596-621, timer1_trainer_ft.py:551-566): video_inputs = [torch.load(video_inputs.pt)] (a list
        holding the list process_vision_info_v3 returned), video_kwargs = [{"fps": [f]}]; tolerate the un-nested forms."""
        v = example["video_inputs"]
        while isinstance(v, (list, tuple)):
            v = v[0]
        kw = example.get("video_kwargs", {"fps": 2.0})
        while isinstance(kw, (list, tuple)):
            kw = kw[0]
        fps = kw.get("fps", 2.0)
        while isinstance(fps, (list, tuple)):
            fps = fps[0]
        return v, float(fps)

    def _prompt_from_example(self, example: dict):
        import torch

        from .engine import Prompt

        if "input_ids" not in example:
            if example.get("video_inputs") is not None:
                example = {**example, **self._host_prompt_inputs(example, *self._unwrap_ft(example))}
            elif self.vision_processor is not None:
                example = {**example, **self.vision_processor(example)}
            else:
                example = {**example, **self._default_vision_processor(example)}
        grid = [tuple(int(v) for v in g) for g in np.asarray(example["video_grid_thw"]).reshape(-1, 3)]
        if example.get("video_frames") is not None:
            # resized frames [T, 3, H, W] (uint8 / fp32, 0..255) as fetch_video_v3 returns them (vision_process.py:440-475):
            # the processor's rescale / normalise / patchify (timer1_trainer.py:547-556) runs on the device
            fr = example["video_frames"]
            fr = fr if isinstance(fr, torch.Tensor) else torch.as_tensor(np.asarray(fr))
            if fr.dtype != torch.uint8:
                fr = fr.to(torch.float32)
            return Prompt(input_ids=np.asarray(example["input_ids"]).reshape(-1), pixel_values=None, grid_thw=grid,
                          second_per_grid_ts=example.get("second_per_grid_ts"),
                          frames=fr.to(device=self.device, non_blocking=True).contiguous())
        pix = example["pixel_values_videos"]
        pix = pix if isinstance(pix, torch.Tensor) else torch.as_tensor(np.asarray(pix))
        pix = pix.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        return Prompt(input_ids=np.asarray(example["input_ids"]).reshape(-1), pixel_values=pix, grid_thw=grid,
                      second_per_grid_ts=example.get("second_per_grid_ts"))

    def _decode(self, ids: np.ndarray) -> list:
        pc = self.processing_class
        if pc is None:
            raise RuntimeError("a processing_class with batch_decode() is required to turn completions into text")
        return pc.batch_decode(ids, skip_special_tokens=True)

    # ------------------------------------------------------------------ per-epoch sample filter
    def filter_samples(self, samples: Optional[list] = None, k: int = 2500, max_new_tokens: int = 1024, batch: int = 4) -> tuple:
        """One filtering pass of TimeRFT with the CURRENT policy (what scripts/posttrain/train_rl_SF.sh:71-111 runs
        between training epochs through evaluate.py -> calc_difficulty.py -> process_data.py --task 0070_all): greedy
        rollout of every sample, difficulty = 100 x IoU, keep 0 < IoU <= 0.7, difficulty-sorted linspace subsample of k.
        Samples are dataset rows with `qid` and `timestamp` (or `solution`) columns; this rank handles rows
        rank, rank + world, ... (evaluate.py --curr_idx / --total_idx).  Returns (selected_rows, annotated_rows)."""
        from .filtering import GreedyFilter

        all_rows = list(self.train_dataset if samples is None else samples)
        # default qid = GLOBAL row index, so that the per-rank shards merge without collisions
        all_rows = [dict(r, qid=r.get("qid", i), timestamp=r.get("timestamp", r.get("solution"))) for i, r in enumerate(all_rows)]
        rows = all_rows[self.rank::self.world]
        pc = self.processing_class
        if pc is None:
            raise RuntimeError("a processing_class with decode() is required to turn completions into text")
        decode = lambda ids: pc.decode(ids, skip_special_tokens=False)
        flt = GreedyFilter(self.engine, self._prompt_from_example, decode, max_new_tokens=max_new_tokens, batch=batch)
        if self.world == 1:
            return flt.run(rows, k)
        # the reference joins every shard's predictions first (calc_difficulty.py:40-66) and takes ONE difficulty-sorted
        # linspace subsample of the merged list (process_data.py:61-80): annotate locally, gather, select globally
        import torch.distributed as dist

        from .filtering import attach_difficulty, select_0070

        gathered = [None] * self.world
        dist.all_gather_object(gathered, flt.difficulties(rows))
        difficulty = {qid: d for part in gathered for qid, d in part.items()}
        annotated = attach_difficulty(all_rows, difficulty)                 # original row order, every rank identical
        return select_0070(annotated, k), annotated

    # ------------------------------------------------------------------ one micro-step
    def _prepare(self, example: dict):
        """(Prompt, reward callback) for one dataset row — reward protocol of timer1_trainer.py:685-698."""
        G = self.num_generations
        prompt = self._prompt_from_example(example)
        prompts = [self.make_conversation_video(example) if "problem" in example and "video_path" in example else example.get("prompt")] * G

        def reward_fn(comp_ids: np.ndarray) -> np.ndarray:
            completions = self._decode(comp_ids)
            if _is_conversational(example):
                completions = [[{"role": "assistant", "content": c}] for c in completions]
            # every dataset column except prompt / completion (timer1_trainer.py:685-694); the tensors this engine's
            # pre-processed layout adds are not dataset columns of the reference and stay out
            cols = {k: [example[k]] * G for k in example if k not in ("prompt", "completion", "input_ids", "pixel_values_videos", "video_frames", "video_grid_thw", "second_per_grid_ts")}
            out = np.zeros((G, len(self.reward_funcs)), np.float32)
            for i, rf in enumerate(self.reward_funcs):
                out[:, i] = np.asarray(rf(prompts=prompts, completions=completions, **cols), np.float32)
            reward_fn.last = (completions, dict(prompts=prompts, **cols))      # THIS group's texts, for its metric functions
            return out

        dr = self._device_reward_path()
        if dr is not None and "solution" in example:
            # registry rewards: ids -> bytes -> parsed timestamps -> IoU / format flags, all on the device
            reward_fn.device = lambda comp: dr(comp, example["solution"], example.get("durations"))
        return prompt, reward_fn

    def _device_reward_path(self):
        """rewards.DeviceRewards when every reward function is one of the registry's, the tokenizer is byte-level BPE with a
        readable vocabulary and no string clean-up (detok.DetokTable.supports) and nothing else needs the decoded strings
        (the fine-tune trainer's metric functions — the base trainer only stores them, timer1_trainer.py:330 — or
        `device_rewards=False` in the arguments); None otherwise -> the host protocol of timer1_trainer.py:642-698."""
        if not hasattr(self, "_device_rewards"):
            self._device_rewards = None
            from .rewards import DeviceRewards

            from .detok import DetokTable

            tok = getattr(self.processing_class, "tokenizer", self.processing_class)
            needs_text = self._logs_metric_funcs and bool(self.metric_funcs)
            if (getattr(self.args, "device_rewards", True) and not needs_text and DeviceRewards.supports(self.reward_funcs)
                    and DetokTable.supports(tok)):
                table = DetokTable.from_tokenizer(tok, self.device, vocab_size=self.model_cfg.vocab_size)
                self._device_rewards = DeviceRewards(self.reward_funcs, table)
        return self._device_rewards

    def compute_loss(self, model, inputs, return_outputs: bool = False, num_items_in_batch=None):
        """rollout + scoring + rewards + advantages + loss + BACKWARD for one prompt group (the backward the reference
        runs in HF Trainer.training_step is part of this call). Returns the (grad-accum-scaled) loss as a 1-element
        device tensor."""
        if return_outputs:
            raise ValueError("The GRPOTrainer does not support returning outputs")
        return self._accum_step([inputs])[0]

    def _accum_step(self, micro_batches: list) -> list:
        """All micro-batches of one optimizer step at once: they see the same weights, so their rollouts are decoded
        together (one weight stream for all of them); scoring / backward stay per prompt-group."""
        for inputs in micro_batches:
            if len(inputs) != 1:
                raise ValueError("per_device_train_batch_size must be 1 (the reference only consumes inputs[0], timer1_trainer.py:548-551)")
        pre = getattr(self, "_prefetched", None)
        self._prefetched = None
        if pre is not None and pre[0] is micro_batches:
            prepared, handle = pre[1], pre[2]
        else:
            prepared, handle = [self._prepare(inputs[0]) for inputs in micro_batches], None
        accum = self.args.gradient_accumulation_steps
        seed = self.run.seed + 7919 * (self.state.global_step * self.world + self.rank)
        outs = self.engine.grpo_multi_step([p for p, _ in prepared], [r for _, r in prepared], seed=seed, loss_scale=1.0 / accum,
                                           vit_prefetch=handle)
        for out, (_, rf) in zip(outs, prepared):
            self._log_metrics(out, getattr(rf, "last", None))
        return [o["loss"] for o in outs]

    def _prefetch_next(self, micro_batches: list):
        """Look-ahead on the dataloader: prepare the NEXT optimizer step's prompts and run their frozen ViT blocks on a
        side stream while this step's optimizer phase (reduce-scatter, AdamW, all-gather) runs."""
        if not self.run.fix_vit:
            return
        prepared = [self._prepare(inputs[0]) for inputs in micro_batches]
        self._prefetched = (micro_batches, prepared, self.engine.prefetch_vit([p for p, _ in prepared]))

    def _log_metrics(self, out: dict, texts=None):
        """Metrics of ONE prompt-group (timer1_trainer.py:740-777); `texts` = (completions, reward kwargs) of that group
        when the host reward protocol decoded them."""
        import torch

        stats = out["stats"]
        rpf = out["rewards_per_func"].mean(0)
        extra = self._extra_local(texts)                    # subclass: per-rank values that are averaged over ranks
        vec = torch.cat([stats[3:4], rpf, out["rewards"].mean().reshape(1), out["reward_std"].mean().reshape(1),
                         stats[1:2], stats[2:3], stats[5:7], torch.tensor(extra, dtype=stats.dtype, device=stats.device)])
        self._local_clip = stats[5:7]                       # this rank's (low, high) clip fractions, before the average
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(vec, op=dist.ReduceOp.AVG)      # one fused all-reduce replaces 6 gather_for_metrics (:740-777)
        v = vec.tolist()
        n = len(self.reward_funcs)
        self._metrics["completion_length"].append(v[0])
        for i, rf in enumerate(self.reward_funcs):
            self._metrics[f"rewards/{rf.__name__}"].append(v[1 + i])
        self._metrics["reward"].append(v[1 + n])
        self._metrics["reward_std"].append(v[2 + n])
        if self.beta != 0.0:
            self._metrics["kl"].append(v[3 + n])
        self._metrics["generation_entropy"].append(v[4 + n])
        self._extra_metrics(v[5 + n], v[6 + n], v[7 + n:])

    def _extra_local(self, texts) -> list:
        return []

    def _extra_metrics(self, clip_low: float, clip_high: float, extra: list):
        pass

    def log(self, logs: dict, start_time: Optional[float] = None) -> None:
        metrics = {k: sum(v) / len(v) for k, v in self._metrics.items()}
        logs = {**logs, **metrics, "step": self.state.global_step}
        self.state.log_history.append(logs)
        if self.state.is_world_process_zero:
            print(json.dumps({k: (round(v, 6) if isinstance(v, float) else v) for k, v in logs.items()}))
        for cb in self.callbacks:
            cb.on_log(self.args, self.state, self.control, logs=logs, trainer=self)
        self._metrics.clear()

    # ------------------------------------------------------------------ loop
    def train(self, resume_from_checkpoint: Optional[str] = None):
        a = self.args
        loader = self.get_train_dataloader(0)
        accum = max(1, a.gradient_accumulation_steps)
        # HF Trainer arithmetic (trainer.py:_inner_training_loop, 4.51.1), from the per-rank loader length that is common to
        # all ranks: update steps per epoch = len // accum (at least 1) for max_steps; an epoch's trailing partial
        # accumulation still performs an optimizer step.
        steps_per_epoch = max(1, len(loader) // accum)
        if getattr(a, "max_steps", -1) and a.max_steps > 0:
            max_steps = a.max_steps
        else:
            max_steps = math.ceil(a.num_train_epochs * steps_per_epoch)
        if self.state.max_steps <= 0:
            self.state.max_steps = max_steps
        start_step = 0
        if resume_from_checkpoint:
            start_step = self._load_checkpoint(resume_from_checkpoint)
            self.state.global_step = start_step
        for cb in self.callbacks:
            cb.on_train_begin(a, self.state, self.control, trainer=self)
        total = self.state.max_steps if getattr(a, "max_steps", -1) <= 0 else a.max_steps
        epoch = 0
        losses = []
        while self.state.global_step < total and not self.control.should_training_stop:
            loader = self.get_train_dataloader(epoch) if epoch > 0 else loader
            n_mb = len(loader)
            micro = loader[0:accum]
            for i in range(0, n_mb, accum):
                losses.extend(self._accum_step(micro))
                nxt = loader[i + accum:i + 2 * accum] if i + accum < n_mb else None
                if nxt and self.state.global_step + 1 < total:
                    self._prefetch_next(nxt)
                done = self.state.global_step - start_step
                lr = a.learning_rate * max(0.0, 1.0 - done / max(1, total - start_step))   # linear decay, no warm-up
                gnorm = self.engine.optimizer_step(lr=lr)
                self.state.global_step += 1
                self.state.epoch = epoch + min(1.0, (i + accum) / max(1, n_mb))
                if self.state.global_step % max(1, a.logging_steps) == 0:
                    loss_val = float(sum(l.item() for l in losses))
                    losses = []
                    self.log({"loss": loss_val, "grad_norm": float(gnorm.item()), "learning_rate": lr, "epoch": self.state.epoch})
                for cb in self.callbacks:
                    cb.on_step_end(a, self.state, self.control, trainer=self)
                if getattr(a, "save_strategy", "steps") == "steps" and a.save_steps and self.state.global_step % a.save_steps == 0:
                    self._save_checkpoint()
                if self.state.global_step >= total or self.control.should_training_stop:
                    break
                micro = nxt
            epoch += 1
            self.state.epoch = float(epoch)
            for cb in self.callbacks:
                cb.on_epoch_end(a, self.state, self.control, trainer=self)
            if getattr(a, "save_strategy", "steps") == "epoch":
                self._save_checkpoint()
        for cb in self.callbacks:
            cb.on_train_end(a, self.state, self.control, trainer=self)
        return self.state

    # ------------------------------------------------------------------ checkpoints (HF-format export)
    def save_model(self, output_dir: Optional[str] = None):
        output_dir = output_dir or self.args.output_dir
        if not self.state.is_world_process_zero:
            return
        os.makedirs(output_dir, exist_ok=True)
        if hasattr(self.engine, "sync_params"):
            self.engine.sync_params()          # the last optimizer step's parameter all-gather may still be in flight
        save_checkpoint_dir(output_dir, self.model_cfg, self.weights)
        # HF Trainer saves processing_class into every checkpoint: scripts/posttrain/train_rl_SF.sh:76-105 feeds checkpoint-N
        # straight into evaluate.py / vLLM (AutoProcessor.from_pretrained) and into the next round as --model_name_or_path
        pc = self.processing_class
        if pc is not None and hasattr(pc, "save_pretrained"):
            pc.save_pretrained(output_dir)

    def _save_checkpoint(self):
        import torch

        d = os.path.join(self.args.output_dir, f"checkpoint-{self.state.global_step}")
        self.save_model(d)
        if self.state.is_world_process_zero:
            with open(os.path.join(d, "trainer_state.json"), "w") as f:
                json.dump({"global_step": self.state.global_step, "epoch": self.state.epoch, "max_steps": self.state.max_steps,
                           "log_history": self.state.log_history}, f)
        if not getattr(self.args, "save_only_model", False):
            os.makedirs(d, exist_ok=True)
            torch.save({"master": self.engine.master, "m": self.engine.adam_m, "v": self.engine.adam_v,
                        "opt_step": self.engine.opt_step, "world": self.world, "rank": self.rank,
                        "buckets": getattr(getattr(self.engine, "plan", None), "bounds", None)},
                       os.path.join(d, f"optimizer_rank{self.rank}.pt"))

    def _load_checkpoint(self, d: str) -> int:
        import torch

        from .weights import param_specs

        shapes = lambda c: [(sp.name, tuple(sp.shape)) for sp in param_specs(c)]
        if shapes(read_checkpoint_config(d)) != shapes(self.model_cfg):
            raise ValueError(f"checkpoint {d} was written for another model configuration (tensor names / shapes differ)")
        self.weights.load_state_dict(read_state_dict(d))      # straight into the live flat buffers: no second copy in HBM
        opt = os.path.join(d, f"optimizer_rank{self.rank}.pt")
        if os.path.exists(opt):
            s = torch.load(opt, map_location=self.device)
            plan = getattr(self.engine, "plan", None)
            if s.get("world", self.world) != self.world or (plan is not None and s.get("buckets") not in (None, plan.bounds)):
                raise ValueError(f"{opt} was written with another world size / bucket layout: the sharded optimizer state does not "
                                 "line up (resume with the same number of ranks, or delete optimizer_rank*.pt to restart AdamW)")
            self.engine.master.copy_(s["master"]); self.engine.adam_m.copy_(s["m"]); self.engine.adam_v.copy_(s["v"])
            self.engine.opt_step = s["opt_step"]
        else:
            self.engine.reset_master_from_params()
        st = os.path.join(d, "trainer_state.json")
        return json.load(open(st)).get("global_step", 0) if os.path.exists(st) else 0

    def push_to_hub(self, *a, **k):
        raise NotImplementedError("no network access in this build; use save_model() and upload the directory")


class TimeR1_Trainer_ft(TimeR1_Trainer):
    """Fine-tune variant (src/time_r1/rl/timer1_trainer_ft.py): dataset rows carry pre-decoded tensors
    (`preprocessed_path`/video_inputs.pt, :551-556) — here rows carry the processed tensors directly — metric functions are
    logged, never added to the reward (:670-691), and clip-ratio metrics are reported (:820-842)."""

    _logs_metric_funcs = True        # metrics/<name> are logged from the decoded completions (:670-691, 790-795)
    _prompt_types = ("v1", "v2", "v3")   # timer1_trainer_ft.py:511-518

    def _extra_local(self, texts) -> list:
        """metrics/<fn> of THIS group's completions (timer1_trainer_ft.py:670-691): same protocol as the reward functions;
        the per-rank means are averaged over ranks in the fused all-reduce (:790-795 gathers, then averages)."""
        if texts is None:
            return []
        comps, kw = texts
        return [float(np.asarray(mf(completions=comps, **kw), np.float32).mean()) for mf in self.metric_funcs]

    def _extra_metrics(self, clip_low: float, clip_high: float, extra: list):
        # :820-842 — mean over ranks of each rank's fraction, plus the smallest low / largest high fraction of any rank
        import torch

        ext = torch.stack([self._local_clip[0], -self._local_clip[1]])
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(ext, op=dist.ReduceOp.MIN)
        low_min, neg_high_max = ext.tolist()
        self._metrics["clip_ratio/low_mean"].append(clip_low)
        self._metrics["clip_ratio/low_min"].append(low_min)
        self._metrics["clip_ratio/high_mean"].append(clip_high)
        self._metrics["clip_ratio/high_max"].append(-neg_high_max)
        self._metrics["clip_ratio/region_mean"].append(clip_low + clip_high)
        for mf, val in zip(self.metric_funcs, extra):
            self._metrics[f"metrics/{mf.__name__}"].append(val)


# ------------------------------------------------------------------------------------------------
# HF-format checkpoint IO (released-checkpoint tensor names; evaluate.py / vLLM can load the export)
# ------------------------------------------------------------------------------------------------
def hf_config_dict(cfg: ModelConfig) -> dict:
    q25 = cfg.is_q25
    vision = {"depth": cfg.vit_depth, "hidden_size": cfg.vit_hidden, "num_heads": cfg.vit_heads, "patch_size": cfg.patch_size,
              "temporal_patch_size": cfg.temporal_patch_size, "spatial_merge_size": cfg.spatial_merge_size,
              "in_channels": cfg.in_channels, "out_hidden_size": cfg.hidden_size}
    if q25:
        vision.update(intermediate_size=cfg.vit_intermediate, window_size=cfg.window_size,
                      fullatt_block_indexes=list(cfg.fullatt_block_indexes), tokens_per_second=cfg.tokens_per_second)
    else:
        ratio = cfg.vit_intermediate / cfg.vit_hidden            # Qwen2VLVisionConfig.mlp_ratio is an int (4 in every release)
        # Qwen2VLVisionConfig: embed_dim = ViT width, hidden_size = the merger's OUTPUT width (the LLM hidden size)
        vision.update(embed_dim=cfg.vit_hidden, hidden_size=cfg.hidden_size, mlp_ratio=int(ratio) if ratio == int(ratio) else ratio)
        vision.pop("out_hidden_size")
    return {
        "model_type": "qwen2_5_vl" if q25 else "qwen2_vl",
        "architectures": ["Qwen2_5_VLForConditionalGeneration" if q25 else "Qwen2VLForConditionalGeneration"],
        "hidden_size": cfg.hidden_size, "intermediate_size": cfg.intermediate_size, "num_hidden_layers": cfg.num_hidden_layers,
        "num_attention_heads": cfg.num_attention_heads, "num_key_value_heads": cfg.num_key_value_heads,
        "vocab_size": cfg.vocab_size, "rms_norm_eps": cfg.rms_norm_eps, "rope_theta": cfg.rope_theta,
        "rope_scaling": {"type": "mrope", "mrope_section": list(cfg.mrope_section)},
        "tie_word_embeddings": cfg.tie_word_embeddings, "torch_dtype": "bfloat16", "vision_config": vision,
        "video_token_id": cfg.video_token_id, "image_token_id": cfg.image_token_id,
        "vision_start_token_id": cfg.vision_start_token_id, "eos_token_id": cfg.eos_token_id, "pad_token_id": cfg.pad_token_id,
    }


def save_checkpoint_dir(path: str, cfg: ModelConfig, weights) -> None:
    import torch

    sd = {k: v.detach().contiguous().cpu() for k, v in weights.export_state_dict().items()}
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(hf_config_dict(cfg), f, indent=1)
    with open(os.path.join(path, "generation_config.json"), "w") as f:      # what the released checkpoints ship (vllm_infer.py:112)
        json.dump({"bos_token_id": cfg.pad_token_id, "pad_token_id": cfg.pad_token_id, "do_sample": True,
                   "eos_token_id": [int(t) for t in cfg.stop_token_ids]}, f, indent=1)
    try:
        from safetensors.torch import save_file

        save_file(sd, os.path.join(path, "model.safetensors"), metadata={"format": "pt"})
    except ImportError:
        torch.save(sd, os.path.join(path, "pytorch_model.bin"))


def read_checkpoint_config(path: str) -> ModelConfig:
    hf = json.load(open(os.path.join(path, "config.json")))
    cfg = from_hf_config(hf)
    for k in ("eos_token_id", "pad_token_id"):
        if isinstance(hf.get(k), int):
            setattr(cfg, k, hf[k])
    cfg.stop_token_ids = (cfg.eos_token_id, cfg.pad_token_id)
    return cfg


def read_state_dict(path: str) -> dict:
    """HF-checkpoint tensors of a directory (safetensors shards or pytorch_model.bin) on the host, released-checkpoint names."""
    import glob

    import torch

    sd: dict = {}
    files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if files:
        from safetensors.torch import load_file

        for f in files:
            sd.update(load_file(f))
    else:
        sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu")
    # transformers >= 4.52 nests the towers ("model.visual.*", "model.language_model.*"): normalise to released names
    return {k.replace("model.visual.", "visual.").replace("model.language_model.", "model."): v for k, v in sd.items()}


def load_checkpoint(path: str, device, world_size: int = 1, train_vit: bool = False):
    from .weights import ModelWeights

    cfg = read_checkpoint_config(path)
    w = ModelWeights(cfg, device, world_size=world_size, train_vit=train_vit)
    w.load_state_dict(read_state_dict(path))
    return cfg, w
