"""One optimizer step of the process-per-GPU trainer: the body of
`Trainer.train_one_iter` (src/Trainer01.py:61-107) for one rank of a ray-sharded job.

    step = RayShardedTrainStep(configs, model, FusedLossComputer(configs, ray_sharded=True), FusedAdam(model.parameters()))
    losses = step(shard_rays(batch, rank, world))          # or step(batch, shard=True)

Like the reference it splits the rank's rays into `configs['sub_batch_size']` sub-batches, runs forward, losses and
backward per sub-batch (gradients accumulate), then applies one optimizer step; unlike `torch.nn.DataParallel`
(Trainer01.py:514) the only exchange is the all-reduce of the gradient buckets, launched from autograd hooks while the
last sub-batch's backward is still running.  Loss values stay on the device (the reference calls `.item()` on every loss
of every sub-batch, ten synchronisations per sub-batch): read them when you need them."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from .distributed import GradientExchange, balanced_bounds


class RayShardedTrainStep:
    def __init__(self, configs: dict, model: torch.nn.Module, loss_computer, optimizer, group=None):
        self.configs, self.model, self.loss_computer, self.optimizer, self.group = configs, model, loss_computer, optimizer, group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.world = world
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        params = [p for p in model.parameters() if p.requires_grad]
        self.exchange = GradientExchange(params, weight=1.0 / world, group=group) if world > 1 else None
        self._equal_checked: set = set()

    def _pieces(self, n: int, sub: int, shard: bool):
        """Ray ranges of this rank, one per backward.  Every rank must run the SAME number of backward passes (each does
        collectives: the loss-count all-reduce and, in the last one, the gradient buckets).
        shard=True : `n` is the GLOBAL batch; the reference's sub-batches [k sub, (k+1) sub) (Trainer01.py:84) are kept and
                     each is split over the ranks, so the count follows from the global sizes alone and every sub-batch is
                     the same set of rays as in the reference (unequal pieces are weighted by their counts).
        shard=False: `n` is this rank's own shard; all ranks must hold equally many rays (checked once per size)."""
        if shard:
            pieces = []
            for start in range(0, n, sub):
                lo, hi = balanced_bounds(min(n, start + sub) - start, self.rank, self.world)
                if hi <= lo:
                    raise ValueError(f'sub-batch of {min(n, start + sub) - start} rays cannot be split over {self.world} ranks')
                pieces.append((start + lo, start + hi))
            return pieces
        if self.world > 1 and n not in self._equal_checked:
            t = torch.tensor([n, -n], dtype=torch.int64, device=next(self.model.parameters()).device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            if int(t[0]) != n or int(-t[1]) != n:
                raise ValueError(f'rank {self.rank} holds {n} rays but the ranks hold between {int(-t[1])} and {int(t[0])}: pass the '
                                 'global batch with shard=True, or equal shards')
            self._equal_checked.add(n)
        return [(s, min(n, s + sub)) for s in range(0, n, sub)]

    def __call__(self, input_batch: Dict, shard: bool = False) -> Dict[str, torch.Tensor]:
        self.optimizer.zero_grad(set_to_none=True)                                               # Trainer01.py:80
        n = input_batch['rays_o'].shape[0]
        sub = self.configs.get('sub_batch_size', n) or n
        iter_losses: Dict[str, torch.Tensor] = {}
        pieces = self._pieces(n, sub, shard)
        for i, (lo, hi) in enumerate(pieces):                                                    # :84-101
            sub_batch = {}
            for key, v in input_batch.items():
                if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n:
                    sub_batch[key] = v[lo:hi]
                elif key == 'common_data':
                    sub_batch[key] = dict(v)
                else:
                    sub_batch[key] = v
            out = self.model(sub_batch)
            losses = self.loss_computer.compute_losses(sub_batch, out)
            if self.exchange is not None:       # the buckets are complete (and may be exchanged) only in the last backward
                self.exchange.arm(i == len(pieces) - 1)
            losses['TotalLoss'].backward()
            for name, value in losses.items():                                                   # update_losses_dict_, num_samples_=1
                value = value['loss_value'] if isinstance(value, dict) else value
                value = value.detach() if isinstance(value, torch.Tensor) else torch.as_tensor(float(value))
                iter_losses[name] = value if name not in iter_losses else iter_losses[name] + value
        if self.exchange is not None:
            self.exchange.finish()
        self.optimizer.step()                                                                    # :102
        return iter_losses
