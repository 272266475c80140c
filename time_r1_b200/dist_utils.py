"""Host-side distributed helpers (device-agnostic; exercised with gloo in the CPU tests).

Multi-GPU model: prompt-groups are independent, rank r takes groups r, r+W, ...; the only exchange per
optimizer step is reduce-scatter(grads) -> local AdamW on the 1/W shard -> all-gather(params), plus one fused
all-reduce of the metric vector."""
from __future__ import annotations

import random


def shard_bounds(numel: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of this rank's slice of the flat trainable buffer (numel is padded to a multiple of 512*world)."""
    assert numel % world == 0, "flat buffer must be padded to a multiple of the world size"
    n = numel // world
    return rank * n, (rank + 1) * n


def rank_indices(n_items: int, world: int, rank: int, seed: int) -> list[int]:
    """This rank's dataset indices for one epoch: a seeded global shuffle, then a strided split."""
    idx = list(range(n_items))
    random.Random(seed).shuffle(idx)
    return idx[rank::world]


def zero2_step(flat_param, flat_grad, shard_update, world: int, rank: int, group=None):
    """reduce-scatter(avg) -> shard_update(grad_shard) -> all-gather. `shard_update(g)` returns the new param shard.
    Works on any backend; the engine calls the NCCL path with its AdamW kernel as `shard_update`."""
    import torch
    import torch.distributed as dist

    lo, hi = shard_bounds(flat_param.numel(), world, rank)
    if world > 1:
        g = torch.empty(hi - lo, dtype=flat_grad.dtype, device=flat_grad.device)
        if dist.get_backend(group) == "gloo":      # gloo has no reduce_scatter_tensor: emulate with all_reduce
            tmp = flat_grad.clone()
            dist.all_reduce(tmp, group=group)
            g.copy_(tmp[lo:hi] / world)
        else:
            dist.reduce_scatter_tensor(g, flat_grad, op=dist.ReduceOp.AVG, group=group)
    else:
        g = flat_grad
    new_shard = shard_update(g)
    if world > 1:
        parts = [torch.empty_like(new_shard) for _ in range(world)]
        dist.all_gather(parts, new_shard.contiguous(), group=group)
        flat_param.copy_(torch.cat(parts))
    else:
        flat_param.copy_(new_shard)
    return g
