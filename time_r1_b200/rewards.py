"""Verifiable rewards of Time-R1 (drop-in for main.py:122-239 / reward_funcs_registry :416-420).

Two routes, same numbers:

* protocol route — the functions below take decoded strings like the reference's (protocol at
  src/time_r1/rl/timer1_trainer.py:685-698; user reward callables are arbitrary Python): regex parsing on the host, the
  IoU arithmetic of the whole group in the `t1_iou_reward` kernel in fp64, bit-identical to CPython's float arithmetic;
* device route — `DeviceRewards`: when every configured reward function is one of this registry's, the trainer never
  decodes on the host: `t1_detokenize` -> `t1_parse_answers` (byte-exact restatement of the regexes and float()) ->
  `t1_iou_reward`, completions and rewards stay in HBM and the step has no device->host sync before the loss.
"""
from __future__ import annotations

import re
from typing import Callable, Optional, Sequence

import numpy as np

_ANSWER_RE = re.compile(r"<answer>(.*?)</answer>", re.DOTALL)
_SPAN_RE = re.compile(r"(\d+\.?\d*) (to|and) (\d+\.?\d*)", re.IGNORECASE)
_FORMAT_RE = re.compile(r"<think>.*?</think>\s*<answer>.*?</answer>", re.DOTALL)


def parse_timestamp_output(text: str) -> Optional[tuple[float, float]]:
    """Last `<number> to|and <number>` inside the last <answer>...</answer> block, or None (main.py:122-142)."""
    blocks = _ANSWER_RE.findall(text)
    if not blocks:
        return None
    spans = _SPAN_RE.findall(blocks[-1])
    if not spans:
        return None
    s, _, e = spans[-1]
    return float(s), float(e)


def _text_of(completion) -> str:
    # conversational datasets wrap completions as [{"role": "assistant", "content": ...}] (timer1_trainer.py:645-649)
    if isinstance(completion, str):
        return completion
    return completion[0]["content"]


def _iou_on_device(mode: int, completions: Sequence, solution: Sequence, durations: Optional[Sequence]) -> list[float]:
    import torch

    from . import ops

    n = len(completions)
    if n == 0:
        return []
    ps, pe, valid = np.zeros(n), np.zeros(n), np.zeros(n, np.uint8)
    for i, c in enumerate(completions):
        parsed = parse_timestamp_output(_text_of(c))
        if parsed is not None:
            ps[i], pe[i], valid[i] = parsed[0], parsed[1], 1
    gt = np.asarray([[float(s[0]), float(s[1])] for s in solution], np.float64).reshape(n, 2)
    dur = np.asarray([float(d) for d in durations], np.float64) if durations is not None else np.ones(n)
    dev = "cuda"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    o64, _ = ops.iou_reward(mode, t(ps), t(pe), t(valid), t(gt[:, 0]), t(gt[:, 1]), t(dur))
    return o64.cpu().tolist()


def iou_timestamp_reward(completions, solution, **kwargs) -> list[float]:
    """IoU of the predicted [start, end] against the ground truth; 0 when unparsable (main.py:145-181)."""
    return _iou_on_device(0, completions, solution, None)


def iou_timestamp_reward_v2(completions, solution, **kwargs) -> list[float]:
    """IoU x (1 - |dS|/dur) x (1 - |dE|/dur) (main.py:184-231)."""
    return _iou_on_device(1, completions, solution, kwargs.get("durations"))


def format_reward(completions, **kwargs) -> list[float]:
    """1.0 iff the stripped completion is exactly <think>...</think>\\s*<answer>...</answer> (main.py:234-239)."""
    return [1.0 if _FORMAT_RE.fullmatch(_text_of(c).strip()) else 0.0 for c in completions]


reward_funcs_registry = {
    "iou": iou_timestamp_reward,
    "iou_v2": iou_timestamp_reward_v2,
    "format": format_reward,
}


class DeviceRewards:
    """rewards_per_func [G, n_funcs] fp32 for one prompt-group without leaving the device.

    Replaces, for registry reward functions only, timer1_trainer.py:642-698: batch_decode(skip_special_tokens=True) ->
    reward_func(completions=..., solution=..., durations=...) -> torch.tensor(..., dtype=float32)."""

    _MODES = {iou_timestamp_reward: 0, iou_timestamp_reward_v2: 1}

    def __init__(self, reward_funcs: Sequence, table):
        if not self.supports(reward_funcs):
            raise ValueError("DeviceRewards only covers iou_timestamp_reward, iou_timestamp_reward_v2 and format_reward")
        self.funcs, self.table = list(reward_funcs), table

    @classmethod
    def supports(cls, reward_funcs: Sequence) -> bool:
        return len(reward_funcs) > 0 and all(f in cls._MODES or f is format_reward for f in reward_funcs)

    def __call__(self, completion_ids, solution, duration=None):
        """completion_ids [G, C] int32 device tensor; solution = (gt_start, gt_end) of the prompt; duration for iou_v2."""
        import torch

        from . import ops

        G = completion_ids.shape[0]
        dev = completion_ids.device
        text, text_len = ops.detokenize(completion_ids, self.table)
        ps, pe, valid, _, fmt32 = ops.parse_answers(text, text_len, 0)
        full = lambda v: torch.full((G,), float(v), device=dev, dtype=torch.float64)
        cols = []
        for f in self.funcs:
            if f is format_reward:
                cols.append(fmt32)
                continue
            mode = self._MODES[f]
            if mode == 1 and duration is None:
                raise TypeError("iou_timestamp_reward_v2 needs the `durations` column (main.py:190)")
            _, o32 = ops.iou_reward(mode, ps, pe, valid, full(solution[0]), full(solution[1]), full(duration if mode == 1 else 1.0))
            cols.append(o32)
        return torch.stack(cols, dim=1).contiguous()


# ---- logging-only metric functions (main.py:240-376; logged as metrics/<name> by timer1_trainer_ft.py:670-691,790-795,
# never added to the reward).  Host string logic like the reference's: they see the decoded completions. -----------------
_THINK_RE = re.compile(r"<think>(.*?)</think>", re.DOTALL)
_TIMESTEP_RE = re.compile(r"<timestep>\s*(\d+\.?\d*)\s+to\s+(\d+\.?\d*)\s*</timestep>", re.IGNORECASE | re.DOTALL)
DEFAULT_STRUCTURE_KEYWORDS = ["analyze", "compare", "deduce", "however", "therefore", "because", "step", "observe", "notice",
                              "identify", "wait"]


def extract_think_content(completion: str) -> Optional[str]:
    """Stripped content of the LAST <think>...</think> block, or None (main.py:240-245)."""
    blocks = _THINK_RE.findall(completion)
    return blocks[-1].strip() if blocks else None


def _think_metric(completions, count: Callable, weight: float, cap: int) -> list[float]:
    """weight x min(count(think), cap) per completion; 0.0 when there is no (or an empty) think block."""
    out = []
    for c in completions:
        think = extract_think_content(_text_of(c))
        out.append(max(0.0, weight * min(count(think), cap)) if think else 0.0)
    return out


def reward_timestep_pair(completions, weight: float = 0.2, max_count: int = 1, **kwargs) -> list[float]:
    """<timestep> a to b </timestep> pairs inside the think block, capped (main.py:248-275)."""
    return _think_metric(completions, lambda t: len(_TIMESTEP_RE.findall(t)), weight, max_count)


def reward_think_length(completions, weight: float = 0.001, max_length: int = 500, **kwargs) -> list[float]:
    """Characters of the think block, capped (main.py:278-299)."""
    return _think_metric(completions, len, weight, max_length)


def reward_keyword_usage(completions, keywords: Optional[Sequence[str]] = None, weight: float = 0.1, max_count: int = 2, **kwargs) -> list[float]:
    """How many of the structure keywords occur in the lower-cased think block, capped (main.py:317-343)."""
    words = DEFAULT_STRUCTURE_KEYWORDS if keywords is None else keywords
    return _think_metric(completions, lambda t: sum(1 for w in words if w in t.lower()), weight, max_count)


def reward_paragraph_structure(completions, weight: float = 0.05, max_paragraphs: int = 2, **kwargs) -> list[float]:
    """Non-blank lines of the think block, capped (main.py:346-366)."""
    return _think_metric(completions, lambda t: sum(1 for p in t.split("\n") if p.strip()), weight, max_paragraphs)


# main.py:422-428 (diversity_reward_func is commented out there: it needs rouge_score)
metric_funcs_registry: dict = {
    "reward_timestep_pair": reward_timestep_pair,
    "reward_think_length": reward_think_length,
    "reward_keyword_usage": reward_keyword_usage,
    "reward_paragraph_structure": reward_paragraph_structure,
}
