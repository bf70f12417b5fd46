"""Ray-sharded data parallelism (replaces torch.nn.DataParallel, reference src/Trainer01.py:514, src/Tester01.py:42).

One process per GPU, weights replicated, rays split into contiguous shards; the only exchange is an all-reduce of the MLP
gradients per optimizer step (NCCL on GPUs; the same code runs over gloo in the CPU tests).  Rendering shards image rows and
needs no collective."""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n units for `rank` (what DataParallel.scatter hands each replica)."""
    per = (n + world - 1) // world
    return min(n, rank * per), min(n, (rank + 1) * per)


def shard_rays(batch: Dict, rank: int, world: int) -> Dict:
    n = batch['rays_o'].shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return {k: (v[lo:hi] if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n else v) for k, v in batch.items()}


def allreduce_gradients(params: Iterable[torch.nn.Parameter], weight: float = 1.0, group=None) -> None:
    """Sum the gradients of all ranks in one flat bucket.  Each rank first scales its gradient by ``weight`` =
    (rays of this rank that enter the mean) / (such rays over all ranks), so that per-rank *mean* losses add up to the
    gradient of the global mean loss; for equal shards weight = 1/world."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    if weight != 1.0:
        flat *= weight
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, group=group)
    off = 0
    for p in params:
        p.grad.copy_(flat[off:off + p.numel()].view_as(p))
        off += p.numel()
