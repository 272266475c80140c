"""Drop-in for the reference's `src/time_r1` package (src/time_r1/__init__.py:1-6): main.py / finetune.py keep their
`from src.time_r1 import TimeR1_Trainer, TimeR1_Trainer_ft` line and get the B200 engine underneath."""
from time_r1_b200.trainer import TimeR1_Trainer, TimeR1_Trainer_ft  # noqa: F401

__all__ = ["TimeR1_Trainer", "TimeR1_Trainer_ft"]
