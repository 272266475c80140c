"""Prompt templates and the conversation structure of the drop-in trainers against the reference's string constants
(tests/golden/prompt_templates.json, extracted with `ast` from src/time_r1/rl/timer1_trainer.py:61-67 and
timer1_trainer_ft.py:59-71 by oracle/gen_golden_prompts.py) — character for character."""
import json
import os
import types

import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prompt_templates.json")


def test_templates_are_the_reference_strings():
    from time_r1_b200 import trainer as T

    g = json.load(open(GOLD))
    assert T.SYSTEM_PROMPT == g["timer1_trainer.py"]["SYSTEM_PROMPT"] == g["timer1_trainer_ft.py"]["SYSTEM_PROMPT"]
    assert T.QUESTION_TEMPLATE_TG_v1 == g["timer1_trainer.py"]["QUESTION_TEMPLATE_TG_v1"] == g["timer1_trainer_ft.py"]["QUESTION_TEMPLATE_TG_v1"]
    for v in ("v1", "v2", "v3"):
        assert T.QUESTION_TEMPLATES[v] == g["timer1_trainer_ft.py"][f"QUESTION_TEMPLATE_TG_{v}"], v


@pytest.mark.parametrize("v", ["v1", "v2", "v3"])
def test_conversation_structure(v):
    """timer1_trainer.py:490-510 / timer1_trainer_ft.py:511-536: one user turn = [text, video] with the clip bounds and the
    pixel budget of the reference (total_pixels 3584*28*28, min_pixels 16*28*28)."""
    from time_r1_b200.trainer import QUESTION_TEMPLATES, TimeR1_Trainer, TimeR1_Trainer_ft

    ex = dict(problem="person opens a door", video_path="/v/a.mp4", video_start=1.5, video_end=None)
    me = types.SimpleNamespace(prompt_type=v, _prompt_types=TimeR1_Trainer_ft._prompt_types)
    conv = TimeR1_Trainer_ft.make_conversation_video(me, ex)
    assert conv == [{"role": "user", "content": [
        {"type": "text", "text": QUESTION_TEMPLATES[v].replace("[EVENT]", "person opens a door")},
        {"type": "video", "video": "/v/a.mp4", "video_start": 1.5, "video_end": None, "total_pixels": 3584 * 28 * 28, "min_pixels": 16 * 28 * 28}]}]
    assert "[EVENT]" not in conv[0]["content"][0]["text"]
    base = types.SimpleNamespace(prompt_type=v, _prompt_types=TimeR1_Trainer._prompt_types)
    if v == "v1":
        assert TimeR1_Trainer.make_conversation_video(base, ex) == conv
    else:                       # the base trainer only knows v1 (the reference would hit an unbound prompt_text)
        with pytest.raises(ValueError):
            TimeR1_Trainer.make_conversation_video(base, ex)
