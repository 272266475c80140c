from time_r1_b200.trainer import TimeR1_Trainer, TimeR1_Trainer_ft  # noqa: F401
