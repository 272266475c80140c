"""ORACLE (test infrastructure only) — CPU restatement of Time-R1's reward functions and filtering arithmetic.

Follows /root/reference/main.py:122-142 (parse), :145-181 (iou), :184-231 (iou_v2), :234-239 (format) and
src/vllm_inference/calc_difficulty.py:10-37 (compute_IoU x 100).  Pinned by tests/golden/rewards_golden.json (outputs
of the unmodified reference functions, oracle/gen_golden_rewards.py) and the 2500 known-answer rows of
dataset/timer1/annotations/train_2k5.json (tests/golden/iou_train_2k5.json).  Only tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs import this.
"""
import re


def parse_timestamp_output(output_string):
    answers = re.findall(r"<answer>(.*?)</answer>", output_string, re.DOTALL)
    if not answers:
        return None
    matches = re.findall(r"(\d+\.?\d*) (to|and) (\d+\.?\d*)", answers[-1], re.IGNORECASE)
    if not matches:
        return None
    return float(matches[-1][0]), float(matches[-1][2])


def iou(pred, gt):
    """Returns None where the reference leaves `iou` unbound (union <= 0, main.py:165-168)."""
    inter = max(0, min(pred[1], gt[1]) - max(pred[0], gt[0]))
    union = max(pred[1], gt[1]) - min(pred[0], gt[0])
    return inter / union if union > 0 else None


def iou_reward(text, gt):
    p = parse_timestamp_output(text)
    if not p:
        return 0.0
    return iou(p, gt)


def iou_v2_reward(text, gt, duration):
    p = parse_timestamp_output(text)
    if not p:
        return 0.0
    v = iou(p, gt)
    if v is None:
        return None
    gs, ge = 1.0 * gt[0] / duration, 1.0 * gt[1] / duration
    ps, pe = 1.0 * p[0] / duration, 1.0 * p[1] / duration
    return v * (1 - abs(gs - ps)) * (1 - abs(ge - pe))


def format_reward(text):
    return 1.0 if re.fullmatch(re.compile(r"<think>.*?</think>\s*<answer>.*?</answer>", re.DOTALL), text.strip()) else 0.0


def difficulty(pred, gt):
    v = iou(pred, gt)
    return 0.0 if v is None else v * 100
