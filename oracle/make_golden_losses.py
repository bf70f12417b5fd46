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


def reprojection_scene(n, seed, h=48, w=64, focal=50.0, plane=4.0, baseline=1.0):
    """Three views of a textured fronto-parallel plane z = -plane (identity rotations, cameras at x = -b, 0, +b), rendered
    analytically, and n random pixels with rays as DataPreprocessor01.get_rays builds them (:351-368)."""
    g = torch.Generator().manual_seed(seed)
    poses = torch.eye(4)[None].repeat(3, 1, 1)
    poses[:, 0, 3] = torch.tensor([-baseline, 0.0, baseline])
    intr = torch.tensor([[focal, 0, w / 2], [0, focal, h / 2], [0, 0, 1]])[None].repeat(3, 1, 1)
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing='ij')
    images = []
    for v in range(3):
        wx = poses[v, 0, 3] + (xs - w / 2) / focal * plane                 # world X, Y of the plane point seen by the pixel
        wy = -(ys - h / 2) / focal * plane
        images.append(torch.stack([0.5 + 0.5 * torch.sin(2.1 * wx + 0.3 * wy), 0.5 + 0.5 * torch.sin(1.3 * wy - 0.7 * wx),
                                   0.5 + 0.25 * torch.sin(0.9 * wx) + 0.25 * torch.cos(1.7 * wy)], -1))
    images = torch.stack(images)
    flat = torch.randperm(3 * h * w, generator=g)[:n]
    view, idx = flat // (h * w), flat % (h * w)
    x, y = (idx % w), (idx // w)
    rays_d = torch.stack([(x - w / 2) / focal, -(y - h / 2) / focal, -torch.ones(n)], -1).float()
    rays_o = poses[view][:, :3, 3].contiguous()
    pixel_id = torch.stack([view, x, y], -1).int()
    return dict(rays_o=rays_o, rays_d=rays_d, pixel_id=pixel_id, common_data=dict(poses=poses, images=images, intrinsics=intr, resolution=(h, w)))


def reprojection_case(n, seed):
    """Model outputs around the true depth: the main model is good on some rays, each other model on others."""
    scene = reprojection_scene(n, seed)
    g = torch.Generator().manual_seed(seed + 100)
    plane = 4.0
    out = {}
    for k, sigma in (('depth_coarse', 0.6), ('depth_fine', 0.15), ('points_augmentation_depth_coarse', 0.8), ('views_augmentation_depth_coarse', 0.4)):
        noise = sigma * torch.randn(n, generator=g) * (torch.rand(n, generator=g) < 0.7)
        out[k] = (plane + noise).requires_grad_()
    inp = dict(scene, iter_num=20000, indices_mask_nerf=torch.rand((n,), generator=g) < 0.8)
    inp['indices_mask_sparse_depth'] = ~inp['indices_mask_nerf']
    return inp, out


REPROJECTION_LOSSES = [dict(name='PointsAugmentationDepthLoss02', iter_weights={'0': 0, '10000': 0.1}, rmse_threshold=0.1, patch_size=[5, 5]),
                       dict(name='ViewsAugmentationDepthLoss02', iter_weights={'0': 0, '10000': 0.1}, rmse_threshold=0.1, patch_size=[5, 5]),
                       dict(name='CoarseFineConsistencyLoss02', iter_weights={'0': 0, '10000': 0.1}, rmse_threshold=0.1, patch_size=[5, 5])]


def reprojection_golden(store):
    cfg = json.load(open('/root/reference/runs/training/train1021/Configs.json'))
    cfg['losses'] = REPROJECTION_LOSSES
    computer = LossComputer(cfg)
    for tag, (n, seed) in {'r': (1500, 21), 's': (64, 22)}.items():
        inp, out = reprojection_case(n, seed)
        ref_in = dict(inp, common_data={k: (v[None] if isinstance(v, torch.Tensor) else v) for k, v in inp['common_data'].items()})
        res = computer.compute_losses(ref_in, out)        # (compute_losses strips the leading replica dimension of common_data)
        res['TotalLoss'].backward()
        for k, v in inp.items():
            if isinstance(v, torch.Tensor):
                store[f'{tag}_in_{k}'] = v.numpy()
        for k, v in inp['common_data'].items():
            store[f'{tag}_common_{k}'] = np.asarray(v.numpy() if isinstance(v, torch.Tensor) else v)
        for k, v in out.items():
            store[f'{tag}_out_{k}'] = v.detach().numpy()
            store[f'{tag}_grad_{k}'] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
        for lc in REPROJECTION_LOSSES:
            store[f"{tag}_loss_{lc['name']}"] = np.float32(float(res[lc['name']]['loss_value'].detach()))
        store[f'{tag}_loss_TotalLoss'] = np.float32(float(res['TotalLoss'].detach()))
        print(tag, {k: float(v) for k, v in store.items() if k.startswith(f'{tag}_loss_')},
              {k: int((store[f'{tag}_grad_{k}'] != 0).sum()) for k in out})


PAIR_LOSSES = [dict(name='PointsAugmentationDepthLoss01', weight=0.3), dict(name='ViewsAugmentationDepthLoss01', weight=0.2),
               dict(name='CoarseFineConsistencyLoss01', weight=0.5), dict(name='DenseDepthMSE01', weight=0.7)]


def pair_golden(store):
    """The plain two-sided depth losses and the dense-depth MSE.  DenseDepthMSE01 slices the fine depth with an attribute it
    never sets (`self.num_rays`, :41): the attribute is supplied here (= the batch size) so that the module can run at all."""
    cfg = json.load(open('/root/reference/runs/training/train1021/Configs.json'))
    cfg['losses'] = PAIR_LOSSES
    computer = LossComputer(cfg)
    n, tag = 333, 'p'
    inp, out = case(n, 31)
    g = torch.Generator().manual_seed(32)
    inp['dense_depth_values'] = 1 + 4 * torch.rand((n, 1), generator=g)
    computer.losses['DenseDepthMSE01'].num_rays = n
    res = computer.compute_losses(dict(inp), out)
    res['TotalLoss'].backward()
    for k, v in inp.items():
        if isinstance(v, torch.Tensor):
            store[f'{tag}_in_{k}'] = v.numpy()
    for k, v in out.items():
        store[f'{tag}_out_{k}'] = v.detach().numpy()
        store[f'{tag}_grad_{k}'] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
    for lc in PAIR_LOSSES:
        store[f"{tag}_loss_{lc['name']}"] = np.float32(float(res[lc['name']]['loss_value'].detach()))
    store[f'{tag}_loss_TotalLoss'] = np.float32(float(res['TotalLoss'].detach()))
    print(tag, {k: float(v) for k, v in store.items() if k.startswith(f'{tag}_loss_')})


VIS_LOSSES = [dict(name='VisibilityLoss01', weight=0.4), dict(name='VisibilityPriorLoss01', weight=0.25)]


def visibility_golden(store):
    """VisibilityLoss01 / VisibilityPriorLoss01 on synthetic model outputs of the visibility head (vanilla model block)."""
    cfg = json.load(open('/root/reference/runs/training/train1021/Configs.json'))
    cfg['losses'] = VIS_LOSSES
    computer = LossComputer(cfg)
    for tag, with_prior in (('v', True), ('w', False)):
        n = 60
        g = torch.Generator().manual_seed(51 + int(with_prior))
        out = {}
        for level, s in (('coarse', 16), ('fine', 24)):
            out[f'raw_visibility_{level}'] = torch.rand((n, s, 1), generator=g).requires_grad_()
            out[f'visibility_{level}'] = torch.rand((n, s), generator=g).requires_grad_()
            out[f'visibility2_{level}'] = torch.rand((n, 2), generator=g).requires_grad_()
            out[f'raw_visibility2_{level}'] = torch.rand((n, s, 2, 1), generator=g)
        out['raw_visibility_coarse'].data[3, 5, 0] = out['visibility_coarse'].data[3, 5]      # |0|: zero gradient on both sides
        inp = {'iter_num': 100, 'num_frames': 3, 'rays_o': torch.zeros(n, 3), 'indices_mask_nerf': torch.rand((n,), generator=g) < 0.7}
        if with_prior:
            inp['visibility_prior_masks'] = (torch.rand((n, 2), generator=g) < 0.5).float()
        res = computer.compute_losses(dict(inp), out)
        res['TotalLoss'].backward()
        for k, v in inp.items():
            if isinstance(v, torch.Tensor):
                store[f'{tag}_in_{k}'] = v.numpy()
        for k, v in out.items():
            if k.startswith('raw_visibility2'):
                continue
            store[f'{tag}_out_{k}'] = v.detach().numpy()
            store[f'{tag}_grad_{k}'] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
        for lc in VIS_LOSSES:
            store[f"{tag}_loss_{lc['name']}"] = np.float32(float(res[lc['name']]['loss_value'].detach()))
        store[f'{tag}_loss_TotalLoss'] = np.float32(float(res['TotalLoss'].detach()))
        print(tag, {k: float(v) for k, v in store.items() if k.startswith(f'{tag}_loss_')})


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
    reprojection_golden(store)
    pair_golden(store)
    visibility_golden(store)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'losses.npz'), **store)
    print('wrote tests/golden/losses.npz', {k: float(v) for k, v in store.items() if k.endswith('TotalLoss')})


if __name__ == '__main__':
    main()
