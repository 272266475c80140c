"""Verifiable rewards of Time-R1 (drop-in for main.py:122-239 / reward_funcs_registry :416-420).

Text parsing stays on the host (user reward callables are arbitrary Python and receive decoded strings, protocol at
src/time_r1/rl/timer1_trainer.py:685-698); the IoU arithmetic of a whole group runs in the `t1_iou_reward` kernel in
fp64, bit-identical to the CPython float arithmetic of the reference.
"""
from __future__ import annotations

import re
from typing import Optional, Sequence

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
# metric functions of timer1_trainer_ft.py are logging-only host regexes (out of scope)
metric_funcs_registry: dict = {}
