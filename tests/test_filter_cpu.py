"""Per-epoch sample filter: the host logic against goldens generated from the UNMODIFIED reference
functions (oracle/gen_golden_filter.py -> tests/golden/filter_golden.json)."""
import json
import math
import os

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "filter_golden.json")))


def test_extract_answer_matches_reference_evaluate():
    from time_r1_b200 import filtering as F

    for rec in GOLD["records"]:
        assert F.extract_answer(rec["output_text"]) == rec["pred"], rec["output_text"]


def test_extract_answer_force_matches_reference():
    from time_r1_b200 import filtering as F

    for rec, want in zip(GOLD["records"], GOLD["extract_answer_force"]):
        assert F.extract_answer_force(rec["output_text"]) == want, rec["output_text"]


def test_selection_0070_matches_reference_process_data():
    from time_r1_b200 import filtering as F

    for case in GOLD["selection"]:
        items = []
        for it in case["items"]:
            it = dict(it)
            if isinstance(it["difficulty"], dict):              # NaN travels as {"nonfinite": "nan"} in the golden file
                it["difficulty"] = float(it["difficulty"]["nonfinite"])
            items.append(it)
        got = [x["qid"] for x in F.select_0070(items, case["k"])]
        assert got == case["selected_qids"], (case["k"], len(items))


def test_get_difficulty_safe():
    from time_r1_b200 import filtering as F

    assert F.get_difficulty_safe({"difficulty": "12.5"}) == 12.5
    assert F.get_difficulty_safe({"difficulty": float("nan")}) is None
    assert F.get_difficulty_safe({"difficulty": float("inf")}) is None
    assert F.get_difficulty_safe({"difficulty": None}) is None
    assert F.get_difficulty_safe({"difficulty": "abc"}) is None
    assert F.get_difficulty_safe({}) is None
    assert math.isclose(F.get_difficulty_safe({"difficulty": 70}), 70.0)
