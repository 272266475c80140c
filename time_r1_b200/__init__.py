"""time_r1_b200 — B200-native GRPO rollout-and-update engine for Time-R1's post-training hot path.

Host code is Python; every numerical step runs in hand-written sm_100a CUDA behind the C-ABI of libtimer1.so
(include/timer1.h).  torch is used for device memory, streams and torch.distributed only.
"""
__version__ = "0.1.0"
