"""ORACLE tooling — generates tests/golden/filter_golden.json from the UNMODIFIED reference functions of the per-epoch
sample filter: `evaluate.extract_answer` (evaluate.py:125-151), `calc_difficulty.{extract_answer_force, calc_difficulty,
load_new_data}` (src/vllm_inference/calc_difficulty.py) and `process_data.process_ddata(task="0070_all")`
(src/utils/process_data.py).  The modules are executed from /root/reference with their absent third-party imports stubbed.
Run in the build container only:   python oracle/gen_golden_filter.py
"""
import importlib.util
import json
import os
import random
import sys
import tempfile
import types
import warnings

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    stub("data"); stub("data.data_loader", load_tvgbench_filter=lambda **k: [])
    stub("src"); stub("src.vllm_inference"); stub("src.vllm_inference.data", build_dataloader=None)
    stub("src.vllm_inference.vllm_infer", vllmWrapper=object)
    calc = load(os.path.join(REF, "src/vllm_inference/calc_difficulty.py"), "ref_calc_difficulty")
    proc = load(os.path.join(REF, "src/utils/process_data.py"), "ref_process_data")
    ev = load(os.path.join(REF, "evaluate.py"), "ref_evaluate")

    texts = [
        "<think>the person sits</think>\n<answer>12.54 to 17.83</answer><|im_end|>",
        "The event happens from 3 to 9 seconds, then 10.5 and 20.25.",
        "<answer>from 5.0 To 7.25</answer>",                      # case-sensitive: 'To' does not match
        "<answer>1. to 2.</answer>",
        "no numbers here",
        "only one number 42",
        "it starts at 12 seconds and ends at 30.5 seconds",
        "<think>frame 3, frame 17</think><answer>between 12s-30s</answer>",
        "<answer>\n7 to 8\n</answer>",                              # multi-line answer block: no DOTALL in evaluate.py
        "10.00  to 20.00",
        "<answer>0 to 0</answer>",
        "1e3 to 2e3",
        "５ to ９",
        "<answer>100.5 and 3</answer> trailing 7 to 2",
        "",
    ]
    rng = random.Random(0)
    for _ in range(40):
        a, b = sorted(rng.uniform(0, 120) for _ in range(2))
        form = rng.choice(["<think>x</think><answer>{:.2f} to {:.2f}</answer>", "{:.1f} and {:.1f}", "from {:.0f}s until {:.0f}s", "<answer>{:.2f}-{:.2f}</answer>"])
        texts.append(form.format(a, b))
    records = []
    for i, t in enumerate(texts):
        S, E = sorted(rng.uniform(0, 120) for _ in range(2))
        if i % 11 == 10:
            S = E = 0.0                                              # empty union when the prediction is also [0, 0]
        records.append({"qid": f"q{i}", "pred": ev.extract_answer(t, "tg"), "target": [S, E], "duration": 150.0, "output_text": t})
    with tempfile.TemporaryDirectory() as d, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with open(os.path.join(d, "part0.jsonl"), "w") as f:
            for r in records:
                f.write(json.dumps(r) + "\n")
        diff = calc.load_new_data(d)
    force = [calc.extract_answer_force(t) for t in texts]

    # selection golden: synthetic difficulties incl. ties, zeros, > 70, NaN and the shipped 2500-row annotation file
    sel_cases = []
    for n, k in ((10, 4), (57, 20), (300, 64), (2500, 700), (8, 20)):
        items = []
        for i in range(n):
            dval = rng.choice([0.0, 100.0, 70.0, 70.000001, float("nan"), None, "12.5"]) if rng.random() < 0.15 else round(rng.uniform(0, 100), rng.choice([1, 3, 6]))
            items.append({"qid": f"s{n}_{i}", "difficulty": dval})
        with tempfile.TemporaryDirectory() as d:
            src = os.path.join(d, "in.json")
            json.dump(items, open(src, "w"))
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                proc.process_ddata(src, os.path.join(d, "out"), "0070_all", k=k)
            outp = os.path.join(d, "out_0070_all.json")
            chosen = [x["qid"] for x in json.load(open(outp))] if os.path.exists(outp) else []
        sel_cases.append({"k": k, "items": items, "selected_qids": chosen})

    def enc(x):
        if isinstance(x, float) and (x != x or x in (float("inf"), float("-inf"))):
            return {"nonfinite": repr(x)}
        return x

    out = {
        "source": "evaluate.py:125-151, src/vllm_inference/calc_difficulty.py:10-66, src/utils/process_data.py:11-80,108-136 (unmodified)",
        "records": records,
        "extract_answer_force": force,
        "difficulty": {q: {"difficulty": enc(float(v["difficulty"])), "pred": v["pred"]} for q, v in diff.items()},
        "selection": sel_cases,
    }
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "filter_golden.json"), "w") as f:
        json.dump(out, f, indent=0, default=enc)
    print("wrote", os.path.join(OUT, "filter_golden.json"), len(records), "records,", len(sel_cases), "selection cases")


if __name__ == "__main__":
    main()
