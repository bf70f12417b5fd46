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

from .distributed import GradientExchange, shard_rays


class RayShardedTrainStep:
    def __init__(self, configs: dict, model: torch.nn.Module, loss_computer, optimizer, group=None):
        self.configs, self.model, self.loss_computer, self.optimizer, self.group = configs, model, loss_computer, optimizer, group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.world = world
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        params = [p for p in model.parameters() if p.requires_grad]
        self.exchange = GradientExchange(params, weight=1.0 / world, group=group) if world > 1 else None

    def __call__(self, input_batch: Dict, shard: bool = False) -> Dict[str, torch.Tensor]:
        if shard:
            input_batch = shard_rays(input_batch, self.rank, self.world)
        self.optimizer.zero_grad(set_to_none=True)                                               # Trainer01.py:80
        n = input_batch['rays_o'].shape[0]
        sub = self.configs.get('sub_batch_size', n) or n
        iter_losses: Dict[str, torch.Tensor] = {}
        for start in range(0, n, sub):                                                           # :84-101
            sub_batch = {}
            for key, v in input_batch.items():
                if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n:
                    sub_batch[key] = v[start:start + sub]
                elif key == 'common_data':
                    sub_batch[key] = dict(v)
                else:
                    sub_batch[key] = v
            out = self.model(sub_batch)
            losses = self.loss_computer.compute_losses(sub_batch, out)
            losses['TotalLoss'].backward()
            for name, value in losses.items():                                                   # update_losses_dict_, num_samples_=1
                value = value['loss_value'] if isinstance(value, dict) else value
                value = value.detach() if isinstance(value, torch.Tensor) else torch.as_tensor(float(value))
                iter_losses[name] = value if name not in iter_losses else iter_losses[name] + value
        if self.exchange is not None:
            self.exchange.finish()
        self.optimizer.step()                                                                    # :102
        return iter_losses
