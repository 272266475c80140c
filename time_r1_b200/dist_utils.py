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


class BucketPlan:
    """ZeRO-2 exchange in buckets (scripts/zero2.yaml:3-8 semantics: gradients reduce-scattered, optimizer state sharded,
    parameters all-gathered) so that the collectives can overlap the compute that produces / consumes them.

    The flat trainable buffer [0, numel) is cut into `n_buckets` contiguous buckets whose length is a multiple of
    `world * align`; inside bucket b rank r owns the r-th 1/world slice.  A rank's optimizer shard (fp32 master / m / v and
    the reduced bf16 gradient) is the concatenation of its slices in bucket order, so AdamW still runs over contiguous
    memory per bucket and the all-gather of bucket b is IN PLACE: the updated slice already sits at its position of the flat
    parameter buffer (send buffer == receive buffer + rank * count).

    The backward finishes the flat buffer from its END towards its start (lm_head first, merger last — weights.py layout),
    so `ready_buckets(done_from)` hands out buckets from the last one backwards as `done_from` (the lowest offset whose
    gradient is final) moves down."""

    def __init__(self, numel: int, world: int, n_buckets: int = 32, align: int = 512):
        assert numel % (world * align) == 0, "flat buffer must be padded to a multiple of world * align"
        self.numel, self.world = numel, world
        if world == 1:
            n_buckets = 1
        unit = world * align
        size = -(-numel // n_buckets // unit) * unit
        self.bounds = [(lo, min(numel, lo + size)) for lo in range(0, numel, size)]
        self.slice_len = [(hi - lo) // world for lo, hi in self.bounds]
        self.shard_off = [0]
        for s in self.slice_len:
            self.shard_off.append(self.shard_off[-1] + s)
        self.shard = self.shard_off[-1]
        assert self.shard * world == numel
        self._next = len(self.bounds) - 1          # next bucket ready_buckets() will hand out

    def __len__(self):
        return len(self.bounds)

    def slice_of(self, b: int, rank: int) -> tuple[int, int]:
        """[lo, hi) inside the FLAT buffer of rank's slice of bucket b."""
        lo = self.bounds[b][0] + rank * self.slice_len[b]
        return lo, lo + self.slice_len[b]

    def shard_range(self, b: int) -> tuple[int, int]:
        """[lo, hi) inside a rank's contiguous SHARD buffers (master / m / v / reduced gradient) of bucket b."""
        return self.shard_off[b], self.shard_off[b + 1]

    def gather_shard(self, flat, rank: int, out):
        """out[shard] <- this rank's slices of `flat` (used to initialise the fp32 master weights)."""
        for b in range(len(self)):
            lo, hi = self.slice_of(b, rank)
            a, z = self.shard_range(b)
            out[a:z].copy_(flat[lo:hi])
        return out

    def reset(self):
        self._next = len(self.bounds) - 1

    def ready_buckets(self, done_from: int) -> list:
        """Buckets that lie entirely at or above `done_from` and were not handed out since reset(), last bucket first."""
        out = []
        while self._next >= 0 and self.bounds[self._next][0] >= done_from:
            out.append(self._next)
            self._next -= 1
        return out

    # ---- collectives (NCCL in the engine; gloo emulation for the CPU tests) ----
    def reduce_scatter(self, b: int, flat_grad, shard_g, rank: int, group=None, async_op: bool = False):
        """shard_g[shard_range(b)] <- mean over ranks of flat_grad[rank's slice of bucket b]."""
        import torch.distributed as dist

        lo, hi = self.bounds[b]
        a, z = self.shard_range(b)
        if dist.get_backend(group) == "gloo":      # gloo has no reduce_scatter_tensor
            tmp = flat_grad[lo:hi].clone()
            dist.all_reduce(tmp, group=group)
            s = self.slice_len[b]
            shard_g[a:z].copy_(tmp[rank * s:(rank + 1) * s] / self.world)
            return None
        return dist.reduce_scatter_tensor(shard_g[a:z], flat_grad[lo:hi], op=dist.ReduceOp.AVG, group=group, async_op=async_op)

    def all_gather(self, b: int, flat_param, rank: int, group=None, async_op: bool = False):
        """flat_param[bucket b] <- every rank's slice, in place (rank's own slice is already where it belongs)."""
        import torch
        import torch.distributed as dist

        lo, hi = self.bounds[b]
        slo, shi = self.slice_of(b, rank)
        if dist.get_backend(group) == "gloo":
            parts = [torch.empty(shi - slo, dtype=flat_param.dtype) for _ in range(self.world)]
            dist.all_gather(parts, flat_param[slo:shi].clone(), group=group)
            flat_param[lo:hi].copy_(torch.cat(parts))
            return None
        return dist.all_gather_into_tensor(flat_param[lo:hi], flat_param[slo:shi], group=group, async_op=async_op)


def balanced_deal(costs, world: int) -> list:
    """Deal the len(costs) = world * k items of ONE global batch to `world` ranks, k items each, so that the per-rank cost sums
    are as even as a greedy longest-processing-time pass makes them.  The step's wall clock is the slowest rank's (every rank
    meets in the gradient reduce-scatter), and prompt-groups differ 2-3x in tokens (TimeRFT clips of 9 s to 20 min): an
    arbitrary deal left ranks waiting 35 % of the step at cfg4.  WHICH rank processes a sample does not change the
    DP-averaged gradient, so this is exact-math.  Returns `order`: a permutation of range(len(costs)) such that rank r takes
    order[r], order[r + world], ... (the strided deal the trainer and the bench already use)."""
    n = len(costs)
    assert n % world == 0, "a global batch holds the same number of items for every rank"
    k = n // world
    load = [0.0] * world
    taken = [[] for _ in range(world)]
    for i in sorted(range(n), key=lambda j: (-float(costs[j]), j)):
        r = min((r for r in range(world) if len(taken[r]) < k), key=lambda r: (load[r], r))
        taken[r].append(i)
        load[r] += float(costs[i])
    order = [0] * n
    for r in range(world):
        for a, i in enumerate(taken[r]):
            order[a * world + r] = i
    return order
