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
        self.param_groups = [{'params': self.params, 'lr': lr, 'betas': tuple(betas), 'eps': eps}]
        # Like torch.optim.Adam the state is created lazily: the reference builds its optimizer BEFORE the model moves to the
        # device (src/Trainer01.py:516 then :58), so at construction time the parameters may still live on the CPU.
        self._exp_avg: List[torch.Tensor] = []
        self._exp_avg_sq: List[torch.Tensor] = []
        self.step_count = 0

    def _state(self):
        if not self._exp_avg or any(m.device != p.device for m, p in zip(self._exp_avg, self.params)):
            for p in self.params:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise ValueError('FusedAdam needs contiguous fp32 CUDA parameters (there is no CPU fallback)')
            old_m, old_v = self._exp_avg, self._exp_avg_sq
            self._exp_avg = [torch.zeros_like(p) if not old_m else old_m[i].to(p.device) for i, p in enumerate(self.params)]
            self._exp_avg_sq = [torch.zeros_like(p) if not old_v else old_v[i].to(p.device) for i, p in enumerate(self.params)]
        return self._exp_avg, self._exp_avg_sq

    @property
    def exp_avg(self) -> List[torch.Tensor]:
        return self._state()[0]

    @property
    def exp_avg_sq(self) -> List[torch.Tensor]:
        return self._state()[1]

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

    # ---- checkpoints: the torch.optim.Adam layout, so that src/Trainer01.py:352-381 can save with either optimizer and
    # resume with the other ({'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [{..., 'params': [i, ...]}]}) ----
    def state_dict(self) -> dict:
        if self.step_count == 0 and not self._exp_avg:        # never stepped: no state yet (the parameters may still be on the CPU)
            return to_torch_adam_state(0, self.params, self.params, self.param_groups[0])
        return to_torch_adam_state(self.step_count, self.exp_avg, self.exp_avg_sq, self.param_groups[0])

    def load_state_dict(self, state: dict) -> None:
        step, exp_avg, exp_avg_sq, group = from_adam_state(state, len(self.params))
        self.step_count = step
        for dst, src in zip(self.exp_avg, exp_avg):
            dst.zero_() if src is None else dst.copy_(src)
        for dst, src in zip(self.exp_avg_sq, exp_avg_sq):
            dst.zero_() if src is None else dst.copy_(src)
        self.param_groups[0].update({k: v for k, v in group.items() if k in ('lr', 'betas', 'eps')})
        self.param_groups[0]['betas'] = tuple(self.param_groups[0]['betas'])


def _torch_adam_group_defaults() -> dict:
    """The param_group keys of torch.optim.Adam in the running torch version (its load_state_dict REPLACES the group dict,
    so every key its step() reads has to be present)."""
    g = dict(torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))]).state_dict()['param_groups'][0])
    g.pop('params', None)
    return g


def to_torch_adam_state(step_count: int, exp_avg, exp_avg_sq, group: dict) -> dict:
    """State of `FusedAdam` in the layout of torch.optim.Adam.state_dict() (one param group, no weight decay, no amsgrad:
    src/Trainer01.py:516).  torch keeps no state for parameters that were never stepped; neither does this."""
    pg = _torch_adam_group_defaults()
    pg.update(lr=float(group['lr']), betas=tuple(group['betas']), eps=float(group['eps']))
    pg['params'] = list(range(len(exp_avg)))
    state = {}
    if step_count > 0:
        for i, (m, v) in enumerate(zip(exp_avg, exp_avg_sq)):
            state[i] = {'step': torch.tensor(float(step_count)), 'exp_avg': m.detach().clone(), 'exp_avg_sq': v.detach().clone()}
    return {'state': state, 'param_groups': [pg]}


def from_adam_state(state: dict, n_params: int):
    """-> (step, exp_avg list, exp_avg_sq list, group dict) from a torch.optim.Adam state_dict or from the flat layout this
    class wrote in round 1 ({'step', 'exp_avg': [...], 'exp_avg_sq': [...], 'param_groups'}).  Entries are None for
    parameters without state."""
    group = dict(state['param_groups'][0])
    if 'state' not in state:                                   # round-1 layout
        return int(state['step']), list(state['exp_avg']), list(state['exp_avg_sq']), group
    ids = group.get('params', list(range(n_params)))
    if len(state['param_groups']) != 1 or len(ids) != n_params:
        raise ValueError(f"optimizer state has {len(state['param_groups'])} group(s) / {len(ids)} parameters, expected 1 / {n_params}")
    exp_avg, exp_avg_sq, steps = [], [], []
    for pid in ids:
        st = state['state'].get(pid)
        exp_avg.append(None if st is None else st['exp_avg'])
        exp_avg_sq.append(None if st is None else st['exp_avg_sq'])
        if st is not None:
            steps.append(int(float(st['step'])))
    if steps and min(steps) != max(steps):
        raise ValueError('FusedAdam keeps one step count: the per-parameter steps of this checkpoint differ')
    return (steps[0] if steps else 0), exp_avg, exp_avg_sq, group
