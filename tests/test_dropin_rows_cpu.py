"""CPU suite: the three dataset-row layouts the drop-in trainers accept, and the `src` namespace package.

  * main.py rows (`video_path` only, main.py:455-468) -> the reference's own reader `src.utils.process_vision_info_v3`
    (timer1_trainer.py:518-534) + chat template + tokenizer, <|video_pad|> expanded on the host, pixels left for the device;
  * finetune.py rows (`video_inputs` / `video_kwargs`, finetune.py:596-621, timer1_trainer_ft.py:551-566);
  * pre-processed rows (`input_ids` + `pixel_values_videos` | `video_frames`).
"""
import collections
import os
import sys
import types

import numpy as np
import pytest
import torch

from test_trainer_host_cpu import make_trainer


class FakeProcessor:
    """apply_chat_template + tokenizer with the Qwen2-VL placeholder layout: <|vision_start|><|video_pad|><|vision_end|>."""
    eos_token_id = 2043

    def __init__(self, cfg):
        self.cfg = cfg
        self.tokenizer = self
        self.seen = None

    def apply_chat_template(self, conv, tokenize=False, add_generation_prompt=True):
        assert not tokenize and add_generation_prompt
        self.seen = conv
        text = [c["text"] for c in conv[0]["content"] if c["type"] == "text"][0]
        return f"<user><vs><vp><ve>{text}<gen>"

    def __call__(self, text, add_special_tokens=False):
        assert add_special_tokens is False
        c = self.cfg
        head, rest = text.split("<vs><vp><ve>")
        ids = [11] + [c.vision_start_token_id, c.video_token_id, c.vision_start_token_id + 1] + [20 + (ord(ch) % 50) for ch in rest[:30]] + [12]
        return {"input_ids": ids}


def trainer_with(cfg, cls=None):
    from time_r1_b200.trainer import GRPOArguments, TimeR1_Trainer

    args = GRPOArguments(num_generations=2, max_completion_length=4, beta=0.0, fix_vit=True)
    tr = make_trainer(cls or TimeR1_Trainer, args, [])
    tr.model_cfg = cfg
    tr.processing_class = FakeProcessor(cfg)
    return tr


def test_ft_rows_are_tokenised_expanded_and_left_as_pixels_for_the_device():
    from time_r1_b200 import config
    from time_r1_b200.trainer import TimeR1_Trainer_ft

    cfg = config.toy()
    tr = trainer_with(cfg, TimeR1_Trainer_ft)
    frames = torch.randint(0, 256, (5, 3, 56, 84)).float()                 # odd frame count: last frame is repeated
    row = dict(problem="person sits", video_path="/v/a.mp4", video_start=None, video_end=None, solution=(1.0, 2.0), durations=9.0,
               video_inputs=[[frames]], video_kwargs=[{"fps": [2.5]}], use_preprocessed=[True])
    p = tr._prompt_from_example(row)
    t, h, w = 3, 4, 6
    assert p.grid_thw == [(t, h, w)] and p.pixel_values is None
    assert p.frames.shape == (6, 3, 56, 84) and torch.equal(p.frames[5], p.frames[4]) and torch.equal(p.frames[:5], frames)
    nv = t * h * w // 4
    ids = p.input_ids
    assert (ids == cfg.video_token_id).sum() == nv
    a = int(np.nonzero(ids == cfg.video_token_id)[0][0])
    assert ids[a - 1] == cfg.vision_start_token_id and (ids[a:a + nv] == cfg.video_token_id).all() and ids[a + nv] == cfg.vision_start_token_id + 1
    assert p.second_per_grid_ts == [cfg.temporal_patch_size / 2.5]
    assert "person sits" in tr.processing_class.seen[0]["content"][0]["text"]
    # un-nested forms are accepted too
    p2 = tr._prompt_from_example(dict(row, video_inputs=frames, video_kwargs={"fps": 2.5}))
    assert np.array_equal(p2.input_ids, ids)
    with pytest.raises(ValueError):
        tr._prompt_from_example(dict(row, video_inputs=[[torch.zeros(4, 3, 50, 84)]]))      # not a multiple of 28


def test_main_py_rows_go_through_the_references_reader_when_importable(monkeypatch):
    from time_r1_b200 import config

    cfg = config.toy()
    tr = trainer_with(cfg)
    calls = {}

    def process_vision_info_v3(conversations, return_video_kwargs=False):
        calls["messages"] = conversations
        assert return_video_kwargs
        return None, [torch.full((4, 3, 28, 56), 7.0)], {"fps": [2.0]}

    ns = types.ModuleType("src.utils")
    ns.process_vision_info_v3 = process_vision_info_v3
    monkeypatch.setitem(sys.modules, "src.utils", ns)
    row = dict(problem="x", video_path="/videos/q.mp4", video_start=None, video_end=None, solution=(1.0, 2.0), durations=9.0)
    p = tr._prompt_from_example(row)
    content = calls["messages"][0][0]["content"][0]
    assert content == {"type": "video", "video": "/videos/q.mp4", "total_pixels": 3584 * 28 * 28, "min_pixels": 16 * 28 * 28}   # :518-530
    assert p.grid_thw == [(2, 2, 4)] and p.second_per_grid_ts == [1.0] and p.frames.shape == (4, 3, 28, 56)
    # without the reference package (and without an injected vision_processor) the error says what to do
    monkeypatch.setitem(sys.modules, "src.utils", None)
    with pytest.raises(NotImplementedError, match="PYTHONPATH"):
        tr._prompt_from_example(row)
    tr.vision_processor = lambda ex: dict(input_ids=np.array([1, cfg.video_token_id, 2]), pixel_values_videos=np.zeros((4, cfg.patch_dim), np.float32),
                                          video_grid_thw=np.array([[1, 2, 2]]), second_per_grid_ts=[1.0])
    assert tr._prompt_from_example(row).pixel_values.shape == (4, cfg.patch_dim)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/utils"), reason="the reference checkout only exists in the build container")
def test_src_is_a_namespace_package_that_does_not_shadow_the_references_src_utils():
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); sys.path.append('/root/reference'); import src.time_r1, src.utils; "
            "assert src.time_r1.TimeR1_Trainer.__module__ == 'time_r1_b200.trainer'; "
            "assert src.utils.__file__.startswith('/root/reference/'); print(src.utils.smart_resize(360, 640, factor=28, min_pixels=12544, max_pixels=602112))" % root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "(364, 644)"            # an 8.9 s clip at 360x640 -> 364x644


def test_repo_src_has_no_init_py():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert not os.path.exists(os.path.join(root, "src", "__init__.py")), "a regular `src` package would shadow the reference's src.utils / src.vllm_inference"
