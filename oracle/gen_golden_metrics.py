"""ORACLE tooling — generates tests/golden/metric_funcs_golden.json: outputs of the UNMODIFIED reference's logging-only
metric functions (`main.py:240-376`: reward_timestep_pair, reward_think_length, reward_keyword_usage,
reward_paragraph_structure) on a fixed set of completions.  Run in the build container only:

    python oracle/gen_golden_metrics.py
"""
import contextlib
import io
import json
import os

from gen_golden_rewards import OUT, import_reference_main

CASES = [
    "<think>the person sits</think>\n<answer>12.54 to 17.83</answer>",
    "<think>First, analyze the clip. <timestep>3 to 7.5</timestep> shows it.\n\nHowever, wait: <TIMESTEP> 10.0   to\n12 </TIMESTEP></think><answer>3 to 12</answer>",
    "<think> <timestep>1 to 2</timestep> </think><think>second block: step by step, because we notice and observe; therefore identify, compare, deduce</think>",
    "<think>\n\n  \n</think><answer>1 to 2</answer>",
    "<think></think>",
    "no think block <timestep>1 to 2</timestep>",
    "<think>unclosed <timestep>1 to 2</timestep>",
    "<think>" + "long reasoning " * 60 + "</think>",
    "<think>line one\nline two\n\nline three\n   \nline four</think>",
    "<think>ANALYZE Step WAIT</think>",
    "<think>multi\nline <timestep>4to 5</timestep> <timestep>4 to5</timestep> <timestep>٣ to ٤</timestep></think>",
    "<think>a</think> tail <think>  b because  </think>",
    "",
]


def main():
    ref = import_reference_main()
    rows = []
    with contextlib.redirect_stdout(io.StringIO()):   # the reference prints inside these functions
        for text in CASES:
            rows.append(dict(text=text,
                             reward_timestep_pair=ref.reward_timestep_pair([text])[0],
                             reward_think_length=ref.reward_think_length([text])[0],
                             reward_keyword_usage=ref.reward_keyword_usage([text])[0],
                             reward_paragraph_structure=ref.reward_paragraph_structure([text])[0]))
        batch = {name: getattr(ref, name)(CASES, prompts=["p"] * len(CASES), solution=[(0, 1)] * len(CASES))
                 for name in ref.metric_funcs_registry}
    assert sorted(batch) == sorted(ref.metric_funcs_registry)
    with open(os.path.join(OUT, "metric_funcs_golden.json"), "w") as f:
        json.dump(dict(registry=list(ref.metric_funcs_registry), rows=rows, batch=batch), f, ensure_ascii=True, indent=0)
    print("wrote", len(rows), "rows")


if __name__ == "__main__":
    main()
