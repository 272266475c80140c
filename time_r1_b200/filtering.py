"""Per-epoch sample filtering of TimeRFT: greedy rollout of every training sample with the current
policy, difficulty = 100 x IoU of the parsed prediction, selection `0 < IoU <= 0.7` with a difficulty-sorted linspace
subsample.  Drop-in for the three stages `scripts/posttrain/train_rl_SF.sh:71-111` chains between training epochs:

  * `evaluate.py --datatype tg --use_vllm_inference` (greedy vLLM generation, `src/vllm_inference/vllm_infer.py:106-118`:
    temperature 0, stop on {151645, 151643}, stop token kept in the text)   -> GreedyFilter.rollout
  * `src/vllm_inference/calc_difficulty.py`                                 -> fill_difficulty
  * `src/utils/process_data.py --task 0070_all`                             -> select_0070

Text parsing and the selection are host logic (as in the reference); the rollout runs on the engine's decode path (prefill +
captured decode graph, temperature 0) and the IoU arithmetic in the `t1_iou_reward` kernel (mode 2, fp64, bit-identical
to the reference's numpy arithmetic).  There is no CPU fallback for either.
"""
from __future__ import annotations

import math
import re
from typing import Callable, Optional, Sequence

import numpy as np

_SPAN_RE = re.compile(r"(\d+\.?\d*) (to|and) (\d+\.?\d*)")
_ANSWER_RE = re.compile(r"<answer>(.*?)</answer>")
_NUMBER_RE = re.compile(r"\d+(?:\.\d+)?")


def extract_answer(output_string: str) -> list:
    """`evaluate.py:125-151` (datatype "tg"): the LAST `<number> to|and <number>` anywhere in the output; failing that,
    the last one inside the first <answer>...</answer> (single line); else [None, None].  Case-sensitive, unlike the
    training reward's parser."""
    matches = _SPAN_RE.findall(output_string)
    if not matches:
        m = _ANSWER_RE.search(output_string)
        if m:
            inner = _SPAN_RE.findall(m.group(1).strip())
            if inner:
                return [float(inner[-1][0]), float(inner[-1][2])]
        return [None, None]
    last = matches[-1]
    try:
        return [float(last[0]), float(last[2])]
    except ValueError:
        return [None, None]


def extract_answer_force(output_string: str) -> list:
    """`calc_difficulty.py:40-46`: the first two numbers of the output, or [None, None]."""
    nums = [float(x) for x in _NUMBER_RE.findall(output_string)[:2]]
    return nums if len(nums) == 2 else [None, None]


def compute_difficulty(preds: Sequence, targets: Sequence) -> list:
    """100 x IoU(pred, target) per sample on the device (`calc_difficulty.py:10-37`); a prediction containing None
    scores 0.0; an empty union gives NaN / inf exactly as numpy does."""
    import torch

    from . import ops

    n = len(preds)
    if n == 0:
        return []
    ps, pe, valid = np.zeros(n), np.zeros(n), np.zeros(n, np.uint8)
    for i, p in enumerate(preds):
        if p is not None and None not in p:
            ps[i], pe[i], valid[i] = float(p[0]), float(p[1]), 1
    gt = np.asarray([[float(t[0]), float(t[1])] for t in targets], np.float64).reshape(n, 2)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to("cuda")
    o64, _ = ops.iou_reward(2, t(ps), t(pe), t(valid), t(gt[:, 0]), t(gt[:, 1]), t(np.ones(n)))
    return o64.cpu().tolist()


def fill_difficulty(records: Sequence[dict]) -> dict:
    """`calc_difficulty.load_new_data` (:49-66): re-parse unparsable predictions with extract_answer_force, then
    difficulty per qid.  records: {"qid", "pred", "target", "output_text"} -> {qid: {"difficulty", "pred"}}."""
    preds = []
    for r in records:
        pred = r["pred"]
        if pred is None or None in pred:
            pred = extract_answer_force(r["output_text"])
        preds.append(pred)
    diff = compute_difficulty(preds, [r["target"] for r in records])
    return {r["qid"]: {"difficulty": d, "pred": p} for r, d, p in zip(records, diff, preds)}


def attach_difficulty(original_data: Sequence[dict], difficulty: dict) -> list:
    """`calc_difficulty.main` (:78-92): annotate the original items, keeping the original order."""
    out = []
    for itm in original_data:
        if itm["qid"] in difficulty:
            itm = dict(itm)
            itm["difficulty"] = difficulty[itm["qid"]]["difficulty"]
            itm["pred"] = difficulty[itm["qid"]]["pred"]
            out.append(itm)
    return out


def get_difficulty_safe(item) -> Optional[float]:
    """`process_data.py:11-24`."""
    d = item.get("difficulty") if isinstance(item, dict) else None
    if d is None:
        return None
    try:
        f = float(d)
    except (ValueError, TypeError):
        return None
    return None if (math.isnan(f) or math.isinf(f)) else f


def difficulty_sorted_sample(items: Sequence[dict], k: int) -> list:
    """`process_data.py:61-80`: stable sort by difficulty descending, then the rows at
    `torch.linspace(0, n-1, k).round().long()` (float32 linspace, round-half-even), de-duplicated."""
    import torch

    if not items or k <= 0:
        return []
    n = len(items)
    actual_k = min(n, k)
    ordered = sorted(items, key=lambda x: x["difficulty_float"], reverse=True)
    if actual_k >= n:
        return ordered
    idx = torch.linspace(0, n - 1, steps=actual_k).round().long()
    idx = torch.unique(torch.clamp(idx, 0, n - 1))
    return [ordered[i] for i in idx.tolist()]


def select_0070(data: Sequence[dict], k: int = 2500) -> list:
    """`process_data.process_ddata(task="0070_all")` (:108-136): items with a finite difficulty and 0 < p <= 0.7
    (p = difficulty / 100), difficulty-sorted linspace subsample of k; returns the original items."""
    valid = []
    for item in data:
        d = get_difficulty_safe(item)
        if isinstance(item, dict) and d is not None:
            valid.append({"difficulty_float": d, "p_value": d / 100.0, "data": item})
    subset = [v for v in valid if 0 < v["p_value"] <= 0.7]
    return [v["data"] for v in difficulty_sorted_sample(subset, k)]


class GreedyFilter:
    """Greedy (temperature 0) rollout of each sample once, `batch` samples per decode batch (evaluate.py --batch_size 4):
    the samples of a batch share every weight read of the decode step, exactly like the prompt-groups of a training step.

    make_prompt(sample) -> engine.Prompt;  decode(ids: np.ndarray[int]) -> str (tokenizer.decode with
    skip_special_tokens=False: vLLM is asked for include_stop_str_in_output / skip_special_tokens=False)."""

    def __init__(self, engine, make_prompt: Callable, decode: Callable, max_new_tokens: int = 1024, batch: int = 4):
        assert 1 <= batch <= 8, "at most 8 prompts per decode batch (t1_decode_attn_cluster)"
        self.engine, self.make_prompt, self.decode = engine, make_prompt, decode
        self.max_new_tokens, self.batch = int(max_new_tokens), int(batch)

    def rollout(self, samples: Sequence[dict]) -> list:
        """-> the records evaluate.py writes to its jsonl (:213-232)."""
        eng = self.engine
        cfg = eng.cfg
        stop = set(int(t) for t in cfg.stop_token_ids)
        records = []
        for a in range(0, len(samples), self.batch):
            chunk = list(samples[a:a + self.batch])
            prompts = [self.make_prompt(s) for s in chunk]
            embs = [eng.vit_forward(p)[0] for p in prompts]
            outs = eng.rollout_multi(prompts, embs, 1, self.max_new_tokens, temperature=0.0, top_k=0, seed=0)
            for s, o in zip(chunk, outs):
                ids = o[0].cpu().numpy().astype(np.int64)
                # keep everything up to and including the first stop token (include_stop_str_in_output=True)
                cut = len(ids)
                for j, t in enumerate(ids):
                    if int(t) in stop:
                        cut = j + 1
                        break
                text = self.decode(ids[:cut])
                records.append({"qid": s["qid"], "pred": extract_answer(text), "target": list(s["timestamp"]),
                                "duration": s.get("duration"), "output_text": text})
        return records

    def difficulties(self, samples: Sequence[dict]) -> dict:
        """{qid: {difficulty, pred}} of these samples (what one evaluate.py shard + calc_difficulty.py produce)."""
        return fill_difficulty(self.rollout(samples))

    def run(self, samples: Sequence[dict], k: int = 2500) -> tuple:
        """One filtering epoch: rollout -> difficulty -> 0 < IoU <= 0.7 selection.  Returns (selected, annotated)."""
        records = self.rollout(samples)
        annotated = attach_difficulty(samples, fill_difficulty(records))
        return select_0070(annotated, k), annotated
