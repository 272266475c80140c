"""ORACLE tooling — generates tests/golden/cli_golden.json: the UNMODIFIED reference `main.py` imported with
`trl` := time_r1_b200.cli and `src.time_r1` := this repository's drop-in package (deepspeed / rouge_score, which only
supply names main.py never calls on this path, are stubbed), then ITS OWN parser line (main.py:629-630)
`TrlParser((GRPOScriptArguments, MY_GRPOConfig, ModelConfig)).parse_args_and_config()` run on the flags of the shipped
launch scripts.  Proves that main.py's dataclass subclassing and flag surface work without trl.  Build container only:

    python oracle/gen_golden_cli.py
"""
import dataclasses
import json
import os
import sys
import types

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")

# scripts/posttrain/train_rl.sh:18-45 (shell variables replaced by literals)
TRAIN_RL = """--deepspeed scripts/zero3_offload.json --output_dir ./logs/run --model_name_or_path ./ckpts/Qwen2.5-VL-7B-Instruct
--train_data_path ./dataset/timer1/annotations/train_2k5.json --dataset_name xxx --max_prompt_length 8192
--max_completion_length 200 --num_generations 8 --per_device_train_batch_size 1 --gradient_accumulation_steps 2
--logging_steps 1 --bf16 --torch_dtype bfloat16 --data_seed 42 --gradient_checkpointing true
--attn_implementation flash_attention_2 --fix_vit true --slide_window false --num_train_epochs 5 --run_name exp
--report_to tensorboard --reward_funcs iou_v2 format --temperature 1.0 --prompt_type v1 --is_curriculum_learning false
--logging_dir ./logs/run --save_steps 50 --save_only_model true""".split()
# scripts/posttrain/train_rl_SF.sh:40-69
TRAIN_RL_SF = """--deepspeed scripts/zero3_offload.json --output_dir ./logs/sf --model_name_or_path ./ckpts/prev
--train_data_path ./data.json --video_folder xxx --dataset_name xxx --max_prompt_length 8192 --max_completion_length 20
--num_generations 8 --per_device_train_batch_size 1 --gradient_accumulation_steps 2 --logging_steps 1 --bf16
--torch_dtype bfloat16 --data_seed 42 --gradient_checkpointing true --attn_implementation flash_attention_2 --fix_vit true
--slide_window false --num_train_epochs 1 --run_name sf --report_to tensorboard --reward_funcs iou_v2 format
--temperature 1.0 --prompt_type v1 --is_curriculum_learning false --logging_dir ./logs/sf/sf --save_strategy epoch
--is_early_stopping true --save_only_model false""".split()


# scripts/finetune/run_charades.sh:17-48 (finetune.py, TimeR1_Trainer_ft)
RUN_CHARADES = """--deepspeed scripts/zero3_offload.json --output_dir ./logs/ft --model_name_or_path ./ckpts/Qwen2.5-VL-3B-Instruct_my
--train_data_path ./dataset/finetune/charades/charades_annotation/train.json --video_folder xxx
--preprocessed_data_path ./dataset/finetune/charades/Charades_preprocessed_data_maxpix_3584 --dataset_name xxx
--max_prompt_length 8192 --max_completion_length 1024 --num_generations 8 --per_device_train_batch_size 1
--gradient_accumulation_steps 2 --logging_steps 1 --bf16 --torch_dtype bfloat16 --data_seed 42 --gradient_checkpointing true
--attn_implementation flash_attention_2 --fix_vit true --slide_window false --num_train_epochs 2 --run_name ft
--report_to tensorboard --reward_funcs iou format --temperature 1.0 --beta 0.0 --prompt_type v1 --is_curriculum_learning false
--logging_dir ./logs/finetune/ft --save_steps 100 --use_grpo false --save_only_model true""".split()


def import_reference_main_with_shim():
    sys.path.insert(0, ROOT)
    from time_r1_b200 import cli

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    stub("deepspeed"); stub("deepspeed.runtime"); stub("deepspeed.runtime.fp16")
    stub("deepspeed.runtime.fp16.loss_scaler", LossScaler=type("LossScaler", (), {}))
    stub("deepspeed.runtime.zero"); stub("deepspeed.runtime.zero.config", ZeroStageEnum=type("ZeroStageEnum", (), {}))
    stub("rouge_score", rouge_scorer=types.SimpleNamespace(RougeScorer=object))
    sys.modules["trl"] = cli
    import src.time_r1  # noqa: F401  this repository's drop-in (ROOT precedes the reference on sys.path)

    sys.path.append(REF)
    import main as ref_main

    assert ref_main.TimeR1_Trainer.__module__.startswith("time_r1_b200"), "main.py must have picked up the drop-in trainer"
    return ref_main


def main():
    ref = import_reference_main_with_shim()
    from trl import ModelConfig, TrlParser

    out = {}
    for name, argv in (("train_rl.sh", TRAIN_RL), ("train_rl_SF.sh", TRAIN_RL_SF)):
        parser = TrlParser((ref.GRPOScriptArguments, ref.MY_GRPOConfig, ModelConfig))       # main.py:629
        script_args, training_args, model_args = parser.parse_args_and_config(argv)
        assert [ref.reward_funcs_registry[f] for f in script_args.reward_funcs]            # main.py:548
        out[name] = dict(argv=argv, script_args=dataclasses.asdict(script_args), training_args=dataclasses.asdict(training_args),
                         model_args=dataclasses.asdict(model_args))
    import finetune as ref_ft                                                               # same shim, finetune.py

    assert ref_ft.TimeR1_Trainer_ft.__module__.startswith("time_r1_b200")
    parser = TrlParser((ref_ft.GRPOScriptArguments, ref_ft.MY_GRPOConfig, ModelConfig))
    script_args, training_args, model_args = parser.parse_args_and_config(RUN_CHARADES)
    out["run_charades.sh"] = dict(argv=RUN_CHARADES, script_args=dataclasses.asdict(script_args),
                                  training_args=dataclasses.asdict(training_args), model_args=dataclasses.asdict(model_args))
    # main.main() itself (main.py:549-585) up to the trainer construction: dataset building from an annotation file, the
    # registry lookups and the constructor call must reach the drop-in with kwargs its signature accepts
    import inspect
    import tempfile

    from time_r1_b200.trainer import TimeR1_Trainer as DropIn

    seen = {}

    class Recorder:
        def __init__(self, **kw):
            inspect.signature(DropIn.__init__).bind(None, **kw)          # TypeError if main.py passes something unknown
            seen.update(kwargs=sorted(kw), n_rows=len(kw["train_dataset"]), row_keys=sorted(kw["train_dataset"][0]),
                        reward_funcs=[f.__name__ for f in kw["reward_funcs"]], metric_funcs=[f.__name__ for f in kw["metric_funcs"]],
                        callbacks=[type(c).__name__ for c in kw["callbacks"]])
            raise KeyboardInterrupt                                       # stop main() here: no CUDA in this container

    ann = [dict(video=os.path.join(REF, "assets/OHOFG.mp4"), timestamp=[1.0, 5.5], sentence="Person opens a door.", duration=30.0,
                video_start=None, video_end=None, qid=i) for i in range(3)]
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as tf:
        json.dump(ann, tf)
    ref.TimeR1_Trainer = Recorder
    parser = TrlParser((ref.GRPOScriptArguments, ref.MY_GRPOConfig, ModelConfig))
    script_args, training_args, model_args = parser.parse_args_and_config(TRAIN_RL_SF + ["--train_data_path", tf.name])
    try:
        ref.main(script_args, training_args, model_args)
    except KeyboardInterrupt:
        pass
    os.unlink(tf.name)
    assert seen and seen["n_rows"] == 3, seen
    out["main_constructor_call"] = seen
    with open(os.path.join(OUT, "cli_golden.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote cli_golden.json;", len(out["train_rl.sh"]["training_args"]), "training fields")


if __name__ == "__main__":
    main()
