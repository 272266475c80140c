"""trl-free argument surface (time_r1_b200/cli.py).  tests/golden/cli_golden.json was produced by the UNMODIFIED reference
main.py imported with `trl` := this module and `src.time_r1` := this repository's drop-in, running its own parser line
(main.py:629-630) on the flags of scripts/posttrain/train_rl.sh:18-45 and train_rl_SF.sh:40-69 (oracle/gen_golden_cli.py).
Here the two dataclasses main.py derives (main.py:44-118) are restated and must parse to the same values."""
import dataclasses
import json
import os
from dataclasses import dataclass, field
from typing import Optional

import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cli_golden.json")


def _reference_style_dataclasses():
    from time_r1_b200 import cli

    @dataclass
    class MY_GRPOConfig(cli.GRPOConfig):                       # main.py:44-72
        fix_vit: bool = field(default=False)
        slide_window: bool = field(default=False)
        max_window_layers: int = field(default=2)
        sliding_window_length: int = field(default=4096)
        prompt_type: str = field(default="v1")
        use_grpo: bool = field(default=False)

    @dataclass
    class GRPOScriptArguments(cli.ScriptArguments):            # main.py:75-118
        reward_funcs: list[str] = field(default_factory=lambda: ["iou", "format"])
        max_pixels: Optional[int] = field(default=12845056)
        min_pixels: Optional[int] = field(default=3136)
        train_data_path: str = field(default="./dataset/finetune/charades/Charades/charades_annotation/train.json")
        eval_data_path: str = field(default="./dataset/finetune/charades/Charades/charades_annotation/val.json")
        video_folder: str = field(default="./dataset/finetune/charades/Charades/Charades_v1")
        is_curriculum_learning: bool = field(default=False)
        is_early_stopping: bool = field(default=False)

    return GRPOScriptArguments, MY_GRPOConfig


def _finetune_style_script_arguments():
    from time_r1_b200 import cli

    @dataclass
    class GRPOScriptArguments(cli.ScriptArguments):            # finetune.py's variant: adds preprocessed_data_path
        reward_funcs: list[str] = field(default_factory=lambda: ["iou", "format"])
        max_pixels: Optional[int] = field(default=12845056)
        min_pixels: Optional[int] = field(default=3136)
        train_data_path: str = field(default="./dataset/finetune/charades/Charades/charades_annotation/train.json")
        eval_data_path: str = field(default="./dataset/finetune/charades/Charades/charades_annotation/val.json")
        video_folder: str = field(default="./dataset/finetune/charades/Charades/Charades_v1")
        preprocessed_data_path: Optional[str] = field(default="")
        is_curriculum_learning: bool = field(default=False)

    return GRPOScriptArguments


def test_finetune_launch_script_flags():
    """scripts/finetune/run_charades.sh:17-48 through finetune.py's own dataclasses (golden) vs the restated ones: only the
    flags the script passes are compared for the script arguments (finetune.py's defaults for the rest are its own)."""
    from time_r1_b200 import cli

    g = json.load(open(GOLD))["run_charades.sh"]
    _, T = _reference_style_dataclasses()
    s, t, m = cli.TrlParser((_finetune_style_script_arguments(), T, cli.ModelConfig)).parse_args_and_config(g["argv"])
    assert dataclasses.asdict(t) == g["training_args"] and dataclasses.asdict(m) == g["model_args"]
    for k in ("train_data_path", "video_folder", "preprocessed_data_path", "dataset_name", "reward_funcs", "is_curriculum_learning"):
        assert getattr(s, k) == g["script_args"][k], k
    assert t.beta == 0.0 and t.use_grpo is False and t.max_completion_length == 1024 and t.save_only_model is True


@pytest.mark.parametrize("script", ["train_rl.sh", "train_rl_SF.sh"])
def test_launch_script_flags_parse_like_the_reference_main(script):
    from time_r1_b200 import cli

    g = json.load(open(GOLD))[script]
    S, T = _reference_style_dataclasses()
    script_args, training_args, model_args = cli.TrlParser((S, T, cli.ModelConfig)).parse_args_and_config(g["argv"])
    assert dataclasses.asdict(script_args) == g["script_args"]
    assert dataclasses.asdict(training_args) == g["training_args"]
    assert dataclasses.asdict(model_args) == g["model_args"]
    # what the trainer then reads (timer1_trainer.py:184-206, 305-393)
    assert training_args.num_generations == 8 and training_args.gradient_accumulation_steps == 2 and training_args.bf16 is True
    assert training_args.fix_vit is True and training_args.temperature == 1.0 and script_args.reward_funcs == ["iou_v2", "format"]
    assert training_args.max_completion_length == (200 if script == "train_rl.sh" else 20)
    assert cli.get_peft_config(model_args) is None


def test_parser_behaviour():
    from time_r1_b200 import cli

    S, T = _reference_style_dataclasses()
    p = cli.TrlParser((S, T, cli.ModelConfig))
    s, t, m = p.parse_args_and_config([])
    assert t.beta == 0.04 and t.learning_rate == 1e-6 and s.reward_funcs == ["iou", "format"] and s.max_pixels == 12845056
    s, t, m = p.parse_args_and_config(["--use_grpo", "--beta", "0", "--max_pixels", "1000", "--bf16", "no", "--learning_rate", "2e-6"])
    assert t.use_grpo is True and t.beta == 0.0 and s.max_pixels == 1000 and t.bf16 is False and t.learning_rate == 2e-6
    with pytest.raises(ValueError, match="not used"):
        p.parse_args_and_config(["--no_such_flag", "1"])
    with pytest.raises(SystemExit):
        p.parse_args_and_config(["--fix_vit", "maybe"])
    with pytest.raises(NotImplementedError):
        cli.get_peft_config(cli.ModelConfig(use_peft=True))


def test_config_file_supplies_defaults(tmp_path):
    from time_r1_b200 import cli

    S, T = _reference_style_dataclasses()
    cfg = tmp_path / "run.yaml"
    cfg.write_text("num_generations: 4\nreward_funcs: [iou]\nfix_vit: true\n")
    s, t, m = cli.TrlParser((S, T, cli.ModelConfig)).parse_args_and_config(["--config", str(cfg), "--num_generations", "6"])
    assert t.num_generations == 6 and s.reward_funcs == ["iou"] and t.fix_vit is True


def test_drop_in_constructor_accepts_what_main_py_passes():
    """oracle/gen_golden_cli.py ran the unmodified main.main() (main.py:549-585) with the drop-in in place of the trainer
    class and recorded the constructor call it makes; the drop-in's signature must bind exactly those kwargs, and the
    dataset rows must carry the columns the reward protocol repeats G times (main.py:455-468)."""
    import inspect

    from time_r1_b200 import rewards
    from time_r1_b200.trainer import TimeR1_Trainer, TimeR1_Trainer_ft

    call = json.load(open(GOLD))["main_constructor_call"]
    for cls in (TimeR1_Trainer, TimeR1_Trainer_ft):
        inspect.signature(cls.__init__).bind(None, **{k: None for k in call["kwargs"]})
    assert call["reward_funcs"] == [rewards.reward_funcs_registry[k].__name__ for k in ("iou_v2", "format")]
    assert call["metric_funcs"] == [f.__name__ for f in rewards.metric_funcs_registry.values()]
    assert {"solution", "durations", "problem", "video_path", "video_start", "video_end", "task_type", "choices", "preprocessed_path"} == set(call["row_keys"])
    assert call["callbacks"] == ["StopAfterNEpochsCallback"]
