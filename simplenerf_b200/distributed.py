"""Ray-sharded data parallelism (replaces torch.nn.DataParallel, reference src/Trainer01.py:514, src/Tester01.py:42).

One process per GPU, weights replicated, rays split into contiguous shards; the only exchange is an all-reduce of the MLP
gradients per optimizer step (NCCL on GPUs; the same code runs over gloo in the CPU tests).  Rendering shards image rows and
needs no collective."""
from __future__ import annotations

from typing import Dict, Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n units for `rank` (what DataParallel.scatter hands each replica)."""
    per = (n + world - 1) // world
    return min(n, rank * per), min(n, (rank + 1) * per)


def balanced_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n units for `rank`, sizes differing by at most one (no rank is empty while n >= world)."""
    return n * rank // world, n * (rank + 1) // world


def shard_rays(batch: Dict, rank: int, world: int) -> Dict:
    n = batch['rays_o'].shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return {k: (v[lo:hi] if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n else v) for k, v in batch.items()}


ALL_RAYS = '__all__'     # pseudo mask of mask_count_weights: every ray of the shard (mask-less mean losses)


def mask_count_weights(masks: Dict[str, torch.Tensor], group=None, n_rays: int = None) -> Dict[str, torch.Tensor]:
    """the reference's losses are means over MASKED subsets of the whole batch (MSE01.py:53-59,
    SparseDepthMSE01.py:58-63), computed on the gathered outputs of all replicas.  With one process per GPU every rank takes
    the mean over its own masked rays; scaling the rank's loss of a mask by  n_rank * world / n_global  makes the AVERAGE
    of the ranks' gradients (what the gradient exchange computes) equal the gradient of the global masked mean.  One
    all-reduce of len(masks) counts per step; every scale is 1 when there is one rank.  -> mask name -> 0-dim tensor.
    `n_rays` (the shard's ray count) adds the entry ALL_RAYS for losses that average over every ray: with unequal shards
    their per-rank means need the same n_rank * world / n_global weight."""
    names = sorted(masks)
    device = masks[names[0]].device if names else None
    counts = [masks[k].sum().to(torch.float32) for k in names]
    if n_rays is not None:
        names = names + [ALL_RAYS]
        counts.append(torch.tensor(float(n_rays), device=device))
    local = torch.stack(counts)
    active = dist.is_initialized() and dist.get_world_size(group) > 1
    if not active:
        return {k: torch.ones((), device=local.device) for k in names}
    world = dist.get_world_size(group)
    total = local.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    scale = torch.where(total > 0, local * world / total.clamp(min=1), torch.zeros_like(local))
    return {k: scale[i] for i, k in enumerate(names)}


def _grad_buckets(params: List[torch.nn.Parameter]):
    """Gradients that are views of one flat bucket (the drop-in's backward hands out one zero-padded fp32 bucket per MLP)
    are exchanged in place; anything else goes through a temporary flat copy.  Returns (flat tensors, loose params)."""
    by_storage: Dict[int, list] = {}
    for p in params:
        by_storage.setdefault(p.grad.untyped_storage().data_ptr(), []).append(p)
    flats, loose = [], []
    for ps in by_storage.values():
        g0 = ps[0].grad
        storage = g0.untyped_storage()
        n_el = storage.nbytes() // 4
        spans = sorted((p.grad.storage_offset(), p.grad.numel()) for p in ps)
        tiled = len(ps) > 1 and all(p.grad.dtype == torch.float32 and p.grad.is_contiguous() for p in ps)
        end = 0
        for off, n in spans:                 # every span starts where the previous one ended, up to 3 floats of padding
            tiled = tiled and end <= off <= end + 3
            end = off + n
        tiled = tiled and end <= n_el <= end + 3
        if tiled:
            flats.append(torch.empty(0, dtype=torch.float32, device=g0.device).set_(storage, 0, (n_el,)))
        else:
            loose.extend(ps)
    return flats, loose


def allreduce_gradients(params: Iterable[torch.nn.Parameter], weight: float = 1.0, group=None) -> None:
    """Sum the gradients of all ranks.  Each rank first scales its gradient by ``weight`` =
    (rays of this rank that enter the mean) / (such rays over all ranks), so that per-rank *mean* losses add up to the
    gradient of the global mean loss; for equal shards weight = 1/world.
    Bucketed gradients (one flat buffer per MLP) are all-reduced in place, one collective each, launched back to back."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return
    flats, loose = _grad_buckets(params)
    packed = torch.cat([p.grad.reshape(-1) for p in loose]) if loose else None
    bufs = flats + ([packed] if packed is not None else [])
    active = dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if active else 1
    # NCCL averages in the collective itself when every rank carries the same weight
    use_avg = active and dist.get_backend(group) == 'nccl' and abs(weight * world - 1.0) < 1e-12
    handles = []
    for b in bufs:
        if weight != 1.0 and not use_avg:
            b *= weight
        if active:
            handles.append(dist.all_reduce(b, op=dist.ReduceOp.AVG if use_avg else dist.ReduceOp.SUM, group=group, async_op=True))
    for h in handles:
        h.wait()
    if packed is not None:
        off = 0
        for p in loose:
            p.grad.copy_(packed[off:off + p.numel()].view_as(p))
            off += p.numel()


class GradientExchange:
    """Overlaps the gradient exchange with the backward pass: the all-reduce of an MLP's gradient bucket is launched from
    an autograd hook as soon as the last gradient of that bucket has been produced, so the exchange of the fine MLP
    (whose backward runs first) travels under the backward kernels of the coarse MLPs.

        exchange = GradientExchange(model.parameters(), weight=1 / world)
        loss.backward(); exchange.finish(); optimizer.step()

    The first step learns which parameters share a bucket (it exchanges everything in `finish`); from the second step on
    every complete bucket is reduced in place the moment it is ready.

    overlap=False: nothing is launched from the hooks; `finish` reduces all buckets in ONE coalesced NCCL launch after the
    backward pass.  The backward kernels are persistent and fill every SM, so an all-reduce that runs beside them takes SMs
    from them for as long as it is resident (round 1, 8 GPUs: +0.24 ms of backward for a ~50 us exchange); 9 MB over
    NVLink 5 after the backward costs less than that."""

    def __init__(self, params: Iterable[torch.nn.Parameter], weight: float = 1.0, group=None, overlap: bool = True):
        self.params = [p for p in params if p.requires_grad]
        self.weight, self.group, self.overlap = weight, group, overlap
        self.bucket_of: Dict[int, int] = {}          # id(param) -> bucket index (learned)
        self.bucket_size: List[int] = []
        self._seen: List[int] = []
        self._handles: list = []
        self._done: set = set()
        self.armed = True        # False while a step's earlier sub-batches accumulate into the buckets (see `arm`)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def _active(self) -> bool:
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _reduce(self, buf: torch.Tensor) -> None:
        world = dist.get_world_size(self.group) if self._active() else 1
        use_avg = self._active() and dist.get_backend(self.group) == 'nccl' and abs(self.weight * world - 1.0) < 1e-12
        if self.weight != 1.0 and not use_avg:
            buf *= self.weight
        if self._active():
            self._handles.append(dist.all_reduce(buf, op=dist.ReduceOp.AVG if use_avg else dist.ReduceOp.SUM, group=self.group,
                                                 async_op=True))

    def arm(self, final: bool) -> None:
        """Gradient accumulation (src/Trainer01.py:84-101 runs `sub_batch_size` slices and calls backward on each): the
        buckets are complete only after the LAST backward of the step.  Call `arm(False)` before every earlier backward
        and `arm(True)` before the last one; the hooks launch collectives only while armed, everything else is exchanged
        by `finish`.  With one backward per step nothing has to be called (armed by default)."""
        self.armed = bool(final)

    def _reduce_many(self, bufs: List[torch.Tensor]) -> None:
        """All buffers in one grouped launch where the backend can (NCCL), else one collective each."""
        if self._active() and len(bufs) > 1 and dist.get_backend(self.group) == 'nccl':
            world = dist.get_world_size(self.group)
            use_avg = abs(self.weight * world - 1.0) < 1e-12
            if self.weight != 1.0 and not use_avg:
                torch._foreach_mul_(bufs, self.weight)
            try:
                with dist._coalescing_manager(group=self.group, async_ops=True) as cm:
                    for b in bufs:
                        dist.all_reduce(b, op=dist.ReduceOp.AVG if use_avg else dist.ReduceOp.SUM, group=self.group)
                self._handles.append(cm)
                return
            except (AttributeError, RuntimeError, ValueError):      # no coalescing in this torch build: fall through
                if self.weight != 1.0 and not use_avg:
                    torch._foreach_mul_(bufs, 1.0 / self.weight)
        for b in bufs:
            self._reduce(b)

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if not self.armed or not self.overlap:
            return
        b = self.bucket_of.get(id(p))
        if b is None:
            return
        self._seen[b] += 1
        if self._seen[b] == self.bucket_size[b]:
            flats, loose = _grad_buckets([q for q in self.params if self.bucket_of.get(id(q)) == b and q.grad is not None])
            if len(flats) == 1 and not loose:          # still one tiled bucket: exchange it now
                self._reduce(flats[0])
                self._done.update(id(q) for q in self.params if self.bucket_of.get(id(q)) == b)

    def finish(self) -> None:
        """Exchange whatever the hooks have not (first step, parameters outside buckets), then wait for everything."""
        rest = [p for p in self.params if p.grad is not None and id(p) not in self._done]
        if rest:
            flats, loose = _grad_buckets(rest)
            packed = torch.cat([p.grad.reshape(-1) for p in loose]) if loose else None
            self._reduce_many(flats + ([packed] if packed is not None else []))
        for h in self._handles:
            h.wait()
        if rest and packed is not None:
            off = 0
            for p in loose:
                p.grad.copy_(packed[off:off + p.numel()].view_as(p))
                off += p.numel()
        if not self.bucket_of:                       # learn the buckets from this step's gradients
            by_storage: Dict[int, list] = {}
            for p in self.params:
                if p.grad is not None:
                    by_storage.setdefault(p.grad.untyped_storage().data_ptr(), []).append(p)
            for ps in by_storage.values():
                if len(ps) > 1:
                    for p in ps:
                        self.bucket_of[id(p)] = len(self.bucket_size)
                    self.bucket_size.append(len(ps))
        self._seen = [0] * len(self.bucket_size)
        self._handles, self._done = [], set()
        self.armed = True

    def close(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
