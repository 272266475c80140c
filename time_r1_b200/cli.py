"""Argument surface of the reference's entry points without `trl` (SURVEY §8b "config surface the shim must parse").

`main.py:38` / `finetune.py` do `from trl import GRPOConfig, ModelConfig, ScriptArguments, TrlParser, get_peft_config`,
subclass the first three (`MY_GRPOConfig`, `GRPOScriptArguments`, main.py:44-118) and call
`TrlParser((GRPOScriptArguments, MY_GRPOConfig, ModelConfig)).parse_args_and_config()` (main.py:629-630).  This module
offers those five names with the same behaviour for the flags the shipped scripts pass
(`scripts/posttrain/train_rl.sh:18-45`, `train_rl_SF.sh:40-69`, `scripts/finetune/run_charades.sh`), so the reference's
`main.py` runs on a box that has this package but not trl:

    import sys, time_r1_b200.cli as cli
    sys.modules.setdefault("trl", cli)          # before `import main`

Parsing follows HfArgumentParser: one `--name` per dataclass field, booleans accept `true/false/yes/no/1/0` or stand alone
(`--bf16`), `list[...]` fields take one or more values, `Optional[T]` parses as T, unknown flags are an error.  A leading
`--config file.yaml` supplies defaults, as TrlParser does.
"""
from __future__ import annotations

import argparse
import dataclasses
import sys
import typing
from dataclasses import dataclass, field
from typing import Optional

from .trainer import GRPOArguments


class GRPOConfig(GRPOArguments):
    """trl.GRPOConfig stand-in: every TrainingArguments / GRPOConfig field the reference's scripts and trainer read."""


GRPOConfig = dataclass(GRPOConfig)


@dataclass
class ScriptArguments:
    """trl.ScriptArguments fields the scripts pass (`--dataset_name xxx`, train_rl.sh:22)."""
    dataset_name: Optional[str] = None
    dataset_config: Optional[str] = None
    dataset_train_split: str = "train"
    dataset_test_split: str = "test"


@dataclass
class ModelConfig:
    """trl.ModelConfig fields main.py reads (`model_args.model_name_or_path`, `.attn_implementation`) or the scripts pass
    (`--torch_dtype bfloat16`).  peft fields are accepted so `get_peft_config` can answer None (LoRA is never enabled)."""
    model_name_or_path: Optional[str] = None
    model_revision: str = "main"
    torch_dtype: Optional[str] = None
    trust_remote_code: bool = False
    attn_implementation: Optional[str] = None
    use_peft: bool = False
    lora_r: int = 16
    lora_alpha: int = 32
    lora_dropout: float = 0.05
    load_in_8bit: bool = False
    load_in_4bit: bool = False


def get_peft_config(model_args) -> None:
    """main.py:584 passes the result as peft_config; the shipped scripts never set --use_peft."""
    if getattr(model_args, "use_peft", False):
        raise NotImplementedError("LoRA/peft is never enabled by the reference scripts and is not built (SURVEY §2.2 T8)")
    return None


def _str2bool(v) -> bool:
    if isinstance(v, bool):
        return v
    s = str(v).lower()
    if s in ("yes", "true", "t", "y", "1"):
        return True
    if s in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError(f"truthy value expected, got {v!r} (true/false, yes/no, 1/0)")


def _unwrap_optional(tp):
    if typing.get_origin(tp) is typing.Union:
        args = [a for a in typing.get_args(tp) if a is not type(None)]
        if len(args) == 1:
            return args[0], True
    return tp, False


class TrlParser:
    """`TrlParser(dataclass_types).parse_args_and_config(args=None)` -> one instance per dataclass type, in order."""

    def __init__(self, dataclass_types):
        if dataclasses.is_dataclass(dataclass_types):
            dataclass_types = (dataclass_types,)
        self.dataclass_types = tuple(dataclass_types)
        self.parser = argparse.ArgumentParser(allow_abbrev=False)
        self.parser.add_argument("--config", default=None, help="YAML file with default values for any of the flags")
        seen = set()
        for dc in self.dataclass_types:
            hints = typing.get_type_hints(dc)
            for f in dataclasses.fields(dc):
                if not f.init or f.name in seen:
                    continue
                seen.add(f.name)
                self._add_field(f, hints.get(f.name, str))

    def _add_field(self, f, tp):
        tp, _ = _unwrap_optional(tp)
        default = f.default if f.default is not dataclasses.MISSING else (
            f.default_factory() if f.default_factory is not dataclasses.MISSING else None)
        kw = dict(dest=f.name, default=default, help=(f.metadata or {}).get("help"))
        origin = typing.get_origin(tp)
        if tp is bool:
            kw.update(type=_str2bool, nargs="?", const=True)
        elif origin in (list, typing.List):
            (elem,) = typing.get_args(tp) or (str,)
            kw.update(type=elem if elem in (int, float, str) else str, nargs="+")
        elif tp in (int, float, str):
            kw.update(type=tp)
        else:
            kw.update(type=str)                       # dicts / enums arrive as strings; none of the scripts pass one
        self.parser.add_argument("--" + f.name, **kw)

    def parse_args_and_config(self, args=None, return_remaining_strings: bool = False):
        argv = list(sys.argv[1:] if args is None else args)
        if "--config" in argv:
            import yaml

            i = argv.index("--config")
            with open(argv[i + 1]) as fh:
                self.parser.set_defaults(**(yaml.safe_load(fh) or {}))
            del argv[i:i + 2]
        ns, rest = self.parser.parse_known_args(argv)
        if rest and not return_remaining_strings:
            raise ValueError(f"Some specified arguments are not used by the TrlParser: {rest}")
        out = []
        for dc in self.dataclass_types:
            names = [f.name for f in dataclasses.fields(dc) if f.init]
            out.append(dc(**{n: getattr(ns, n) for n in names}))
        if return_remaining_strings:
            out.append(rest)
        return tuple(out)

    parse_args_into_dataclasses = parse_args_and_config
