"""Logging-only metric functions of the fine-tune trainer (main.py:240-376, registry :422-428) against outputs of the
unmodified reference (oracle/gen_golden_metrics.py -> tests/golden/metric_funcs_golden.json).  Exact float equality."""
import json
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metric_funcs_golden.json")


def test_metric_functions_match_reference_outputs():
    from time_r1_b200 import rewards

    g = json.load(open(GOLD))
    assert list(rewards.metric_funcs_registry) == g["registry"]
    texts = [r["text"] for r in g["rows"]]
    for name, fn in rewards.metric_funcs_registry.items():
        assert fn.__name__ == name                                            # logged as metrics/<__name__>
        for r in g["rows"]:
            assert fn([r["text"]])[0] == r[name], (name, r["text"][:60])
        # batch call with the reward-protocol kwargs the trainer passes (timer1_trainer_ft.py:673-691)
        assert fn(texts, prompts=["p"] * len(texts), solution=[(0, 1)] * len(texts)) == g["batch"][name]
        conv = [[{"role": "assistant", "content": t}] for t in texts]
        assert fn(completions=conv) == g["batch"][name]


def test_metric_knobs():
    from time_r1_b200 import rewards

    t = ["<think>step one\nstep two\nstep three <timestep>1 to 2</timestep><timestep>2 to 3</timestep></think>"]
    assert rewards.reward_timestep_pair(t, max_count=5) == [0.2 * 2]
    assert rewards.reward_paragraph_structure(t, max_paragraphs=10) == [0.05 * 3]
    assert rewards.reward_keyword_usage(t, keywords=["one", "two", "nine"], max_count=9) == [0.1 * 2]
    assert rewards.extract_think_content("<think> a </think><think> b </think>") == "b"
