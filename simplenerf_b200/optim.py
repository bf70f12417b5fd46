"""Optimizer step of the process-per-GPU trainer.

`FusedAdam` performs the update of torch.optim.Adam as the reference configures it (src/Trainer01.py:516: Adam, no weight
decay, no amsgrad) for every parameter in ONE launch of `snerf_adam_step` per 64 tensors.  It keeps the `param_groups` /
`state_dict` surface the reference's LR schedule and checkpointing touch."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List

import torch

from . import _lib, ops

_MAX = 64


class FusedAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        for p in self.params:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise ValueError('FusedAdam needs contiguous fp32 CUDA parameters (there is no CPU fallback)')
        self.param_groups = [{'params': self.params, 'lr': lr, 'betas': tuple(betas), 'eps': eps}]
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.step_count = 0

    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self) -> None:
        group = self.param_groups[0]
        live = [(p, m, v) for p, m, v in zip(self.params, self.exp_avg, self.exp_avg_sq) if p.grad is not None]
        if not live:
            return
        self.step_count += 1
        lib = _lib.load()
        stream = torch.cuda.current_stream(live[0][0].device).cuda_stream
        for i in range(0, len(live), _MAX):
            part = live[i:i + _MAX]
            n = len(part)
            for p, _, _ in part:
                g = p.grad
                if not (g.is_contiguous() and g.dtype == torch.float32 and g.data_ptr() % 16 == 0):
                    p.grad = g.contiguous().clone()
            ptrs = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])          # noqa: E731
            numel = (C.c_longlong * n)(*[p.numel() for p, _, _ in part])
            _lib.check(lib.snerf_adam_step(ptrs([p for p, _, _ in part]), ptrs([p.grad for p, _, _ in part]),
                                           ptrs([m for _, m, _ in part]), ptrs([v for _, _, v in part]), numel, n,
                                           float(group['lr']), float(group['betas'][0]), float(group['betas'][1]),
                                           float(group['eps']), self.step_count, stream), 'snerf_adam_step')
            ops.LAUNCHES['count'] += 1

    def state_dict(self) -> dict:
        return {'step': self.step_count, 'exp_avg': [t.clone() for t in self.exp_avg], 'exp_avg_sq': [t.clone() for t in self.exp_avg_sq],
                'param_groups': [{k: v for k, v in self.param_groups[0].items() if k != 'params'}]}

    def load_state_dict(self, state: dict) -> None:
        self.step_count = int(state['step'])
        for dst, src in zip(self.exp_avg, state['exp_avg']):
            dst.copy_(src)
        for dst, src in zip(self.exp_avg_sq, state['exp_avg_sq']):
            dst.copy_(src)
        self.param_groups[0].update(state['param_groups'][0])
