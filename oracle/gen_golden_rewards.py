"""ORACLE tooling — generates tests/golden/rewards_golden.json and tests/golden/iou_train_2k5.json by importing the
UNMODIFIED reference (`/root/reference/main.py`, with its missing third-party imports stubbed in sys.modules) and
reading its shipped annotation file.  Run in the build container only (the reference does not exist on the GPU box):

    python oracle/gen_golden_rewards.py
"""
import contextlib
import io
import json
import os
import sys
import types
from dataclasses import dataclass

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference_main():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    @dataclass
    class _Cfg:
        pass

    stub("deepspeed"); stub("deepspeed.runtime"); stub("deepspeed.runtime.fp16")
    stub("deepspeed.runtime.fp16.loss_scaler", LossScaler=type("LossScaler", (), {}))
    stub("deepspeed.runtime.zero"); stub("deepspeed.runtime.zero.config", ZeroStageEnum=type("ZeroStageEnum", (), {}))
    stub("rouge_score", rouge_scorer=types.SimpleNamespace(RougeScorer=object))
    stub("src"); stub("src.time_r1", TimeR1_Trainer=object)
    stub("trl", GRPOConfig=_Cfg, ModelConfig=_Cfg, ScriptArguments=_Cfg, TrlParser=object, get_peft_config=lambda *a, **k: None)
    sys.path.insert(0, REF)
    import main as ref_main  # noqa

    return ref_main


CASES = [
    "<think>the person sits</think>\n<answer>12.54 to 17.83</answer>",
    "<think>x</think><answer>0 to 5</answer>",
    "<think>x</think> <answer>3.5 and 9</answer>",
    "<answer>1 to 2</answer><answer>20 to 10</answer>",
    "<think>a</think><answer>from 5.0 to 7.25 seconds</answer> trailing",
    "no tags at all 3 to 4",
    "<think>t</think>\n\n<answer>The event happens in 10.00 TO 20.00.</answer>",
    "<think>t</think><answer>10.00  to 20.00</answer>",          # two spaces: no match
    "<think>t</think><answer>1. to 2.</answer>",
    "<think>t</think><answer>1 to 2 and 3</answer>",
    "<think>t</think><answer>1 to 2, 30 and 40</answer>",
    "<think>\nmulti\nline</think>\n<answer>\n7 to 8\n</answer>",
    "  <think>pad</think><answer>100 to 200</answer>  ",
    "<think></think><answer></answer>",
    "<think>a</think><think>b</think><answer>2 to 3</answer>",
    "<answer>2 to 3</answer><think>late</think>",
    "<think>t</think><answer>５ to ９</answer>",                    # full-width digits: \d is Unicode-aware
    "<think>t</think><answer>1e3 to 2e3</answer>",
    "<think>t</think><answer>-5 to 8</answer>",
    "<think>t</think> <answer>4 to 6</answer>",                # NBSP is \s
    "<think>t</think><answer>0.0 to 0.0</answer>",
    "<think>t</think><answer>12 to 12</answer>",
    "",
]
SOLUTIONS = [(10.0, 20.0), (0.0, 5.0), (1.0, 7.5), (12.0, 12.0), (100.0, 180.0)]
DURATIONS = [30.0, 35.04, 8.9, 1154.4]


def main():
    ref = import_reference_main()
    rows = []
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):   # the reference prints inside its reward functions
        for text in CASES:
            parsed = ref.parse_timestamp_output(text)
            fmt = ref.format_reward([text])[0]
            for sol in SOLUTIONS:
                for dur in DURATIONS:
                    try:
                        iou = ref.iou_timestamp_reward([text], [list(sol)])[0]
                        iou2 = ref.iou_timestamp_reward_v2([text], [list(sol)], durations=[dur])[0]
                    except UnboundLocalError:   # union <= 0 leaves `iou` unbound in the reference (main.py:165-168)
                        iou = iou2 = None
                    rows.append(dict(text=text, parsed=parsed, format=fmt, solution=sol, duration=dur, iou=iou, iou_v2=iou2))
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "rewards_golden.json"), "w") as f:
        json.dump(rows, f, ensure_ascii=True, indent=0)
    ann = json.load(open(os.path.join(REF, "dataset/timer1/annotations/train_2k5.json")))
    ious = [dict(pred=r["pred"], gt=r["timestamp"], duration=r["duration"], difficulty=r["difficulty"]) for r in ann]
    with open(os.path.join(OUT, "iou_train_2k5.json"), "w") as f:
        json.dump(ious, f)
    # calc_difficulty.compute_IoU (numpy) on the same rows, as a second known-answer column
    print(f"wrote {len(rows)} reward rows and {len(ious)} IoU rows to {OUT}")


if __name__ == "__main__":
    main()
