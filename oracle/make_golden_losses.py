"""Generates tests/golden/losses.npz from the UNMODIFIED reference loss modules (/root/reference/src/loss_functions:
LossComputer01, MSE01-03, SparseDepthMSE01-03) on seeded synthetic model outputs.  Run in the build container only:
    python oracle/make_golden_losses.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, '/root/reference/src')
from loss_functions.LossComputer01 import LossComputer  # noqa: E402

OUT_KEYS = ['rgb_coarse', 'rgb_fine', 'points_augmentation_rgb_coarse', 'views_augmentation_rgb_coarse', 'depth_coarse',
            'depth_fine', 'points_augmentation_depth_coarse', 'views_augmentation_depth_coarse']


def case(n, seed, empty_sparse=False):
    g = torch.Generator().manual_seed(seed)
    out = {k: (torch.rand((n, 3), generator=g) if 'rgb' in k else 1 + 4 * torch.rand((n,), generator=g)).requires_grad_() for k in OUT_KEYS}
    inp = {'iter_num': 100, 'rays_o': torch.zeros(n, 3), 'target_rgb': torch.rand((n, 3), generator=g),
           'sparse_depth_values': 1 + 4 * torch.rand((n, 1), generator=g),
           'indices_mask_nerf': torch.rand((n,), generator=g) < 0.75}
    inp['indices_mask_sparse_depth'] = torch.zeros(n, dtype=torch.bool) if empty_sparse else ~inp['indices_mask_nerf']
    return inp, out


def main():
    cfg = json.load(open('/root/reference/runs/training/train1021/Configs.json'))
    cfg['losses'] = [lc for lc in cfg['losses'] if 'MSE' in lc['name']]       # the six masked means of the shipped config
    computer = LossComputer(cfg)
    names = [lc['name'] for lc in cfg['losses']]
    store = {'loss_weights': np.array([lc['weight'] for lc in cfg['losses']], np.float32)}
    for tag, (n, seed, empty) in {'a': (1024, 11, False), 'b': (37, 12, False), 'c': (256, 13, True)}.items():
        inp, out = case(n, seed, empty)
        res = computer.compute_losses(dict(inp), out)
        res['TotalLoss'].backward()
        for k, v in inp.items():
            if isinstance(v, torch.Tensor):
                store[f'{tag}_in_{k}'] = v.numpy()
        for k, v in out.items():
            store[f'{tag}_out_{k}'] = v.detach().numpy()
            store[f'{tag}_grad_{k}'] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
        for name in names:
            store[f'{tag}_loss_{name}'] = np.float32(float(res[name]['loss_value'].detach()))
        store[f'{tag}_loss_TotalLoss'] = np.float32(float(res['TotalLoss'].detach()))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'losses.npz'), **store)
    print('wrote tests/golden/losses.npz', {k: float(v) for k, v in store.items() if k.endswith('TotalLoss')})


if __name__ == '__main__':
    main()
