"""Synthetic LLFF / RealEstate-10K shaped inputs for tests and benchmarks.

Camera constants are the ones the reference froze in
``runs/training/train1021/fern/ModelConfigs.json`` (LLFF, 756x1008) and
``runs/training/train0021/00000/ModelConfigs.json`` (RE10K, 576x1024); the model dictionaries
are the ``model`` block of ``runs/training/train1021/Configs.json``.  Ray construction restates
``DataPreprocessor01.get_rays`` (:351-368), ``get_ndc_rays`` (:371-389) and ``get_view_dirs``
(:392-394) of the reference in numpy fp32.  Nothing here reads ``/root/reference`` at run time.
"""
from __future__ import annotations

import copy
from typing import Dict, Tuple

import numpy as np
import torch

CAMERAS = {
    'llff': dict(resolution=(756, 1008), focal=815.131591796875, centre=(504.0, 378.0),
                 near=1.0, far=6.179403816938637, translation_scale=0.07849926897402267),
    're10k': dict(resolution=(576, 1024), focal=493.9102478027344, centre=(512.0, 288.0),
                  near=1.0, far=133.33334350585938, translation_scale=1.3333333333333333),
}


def _main_mlp(num_samples: int) -> dict:
    return dict(num_samples=num_samples, points_net_depth=8, views_net_depth=1, points_net_width=256,
                views_net_width=128, points_positional_encoding_degree=10,
                views_positional_encoding_degree=4, use_view_dirs=True, view_dependent_rgb=True,
                predict_visibility=False)


def make_configs(kind: str = 'simplenerf', ndc: bool = True, device=None) -> dict:
    """``kind``: 'simplenerf' (4 MLPs, the shipped train1021 model block), 'vanilla' (coarse + fine only: the same
    block with the two augmentation keys removed) or 'simplenerf_fineaug' (6 MLPs: the augmentation models also at
    the fine level, src/models/SimpleNeRF01.py:234-263 -- no shipped config enables them)."""
    model = dict(name='FusedSimpleNeRF01', coarse_mlp=_main_mlp(64), fine_mlp=_main_mlp(128),
                 chunk=4096, lindisp=False, netchunk=16384, perturb=True, raw_noise_std=1.0,
                 white_bkgd=False)
    if kind in ('simplenerf', 'simplenerf_fineaug'):
        pa = _main_mlp(64)
        del pa['num_samples']
        pa['points_sigma_positional_encoding_degree'] = 3
        va = _main_mlp(64)
        del va['num_samples'], va['views_positional_encoding_degree']
        va['use_view_dirs'] = False
        va['view_dependent_rgb'] = False
        model['points_augmentation'] = dict(coarse_mlp=pa)
        model['views_augmentation'] = dict(coarse_mlp=va)
        if kind == 'simplenerf_fineaug':
            model['points_augmentation']['fine_mlp'] = dict(pa)
            model['views_augmentation']['fine_mlp'] = dict(va)
    elif kind != 'vanilla':
        raise ValueError(kind)
    return dict(data_loader=dict(ndc=ndc), model=model, device=[0] if device is None else device)


def _pixel_rays(h: int, w: int, focal: float, cx: float, cy: float, pose: np.ndarray,
                pix: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """pix [n,2] = (x, y) pixel coordinates.  Camera looks down -z, y up (y and z flipped)."""
    x = pix[:, 0].astype(np.float32)
    y = pix[:, 1].astype(np.float32)
    kinv = np.linalg.inv(np.array([[focal, 0, cx], [0, focal, cy], [0, 0, 1]], dtype=np.float32))
    dirs = (kinv[None] @ np.stack([x, y, np.ones_like(x)], 1)[:, :, None])[:, :, 0]
    dirs[:, 1:] *= -1
    rays_d = np.sum(dirs[:, None, :] * pose[:3, :3], -1).astype(np.float32)
    rays_o = np.broadcast_to(pose[:3, -1].astype(np.float32), rays_d.shape).copy()
    return rays_o, rays_d


def _to_ndc(rays_o, rays_d, h, w, focal, near):
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1. / (w / (2. * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1. / (h / (2. * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (w / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1. / (h / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2. * near / rays_o[..., 2]
    return (np.stack([o0, o1, o2], -1).astype(np.float32), np.stack([d0, d1, d2], -1).astype(np.float32))


def make_ray_batch(camera: str, num_rays: int, seed: int, frame: bool = False,
                   start: int = 0) -> Dict[str, torch.Tensor]:
    """The hot path's input dict.  ``frame=False``: ``num_rays`` random pixels
    of 3 views (identity rotation, translations -0.5/0/+0.5 * translation_scale along x), like a
    training batch.  ``frame=True``: rays ``start .. start+num_rays`` of the centre view in
    row-major order, like ``create_test_data``."""
    cam = CAMERAS[camera]
    h, w = cam['resolution']
    focal, (cx, cy) = cam['focal'], cam['centre']
    if frame:
        idx = np.arange(start, start + num_rays, dtype=np.int64)
        view = np.ones_like(idx)
    else:
        g = torch.Generator().manual_seed(seed)
        flat = torch.randperm(3 * h * w, generator=g)[:num_rays].numpy()
        view, idx = flat // (h * w), flat % (h * w)
    pix = np.stack([idx % w, idx // w], 1)
    rays_o = np.zeros((num_rays, 3), np.float32)
    rays_d = np.zeros((num_rays, 3), np.float32)
    for v in range(3):
        pose = np.eye(4, dtype=np.float32)
        pose[0, 3] = (v - 1) * 0.5 * cam['translation_scale']
        sel = view == v
        if sel.any():
            rays_o[sel], rays_d[sel] = _pixel_rays(h, w, focal, cx, cy, pose, pix[sel])
    o_ndc, d_ndc = _to_ndc(rays_o, rays_d, h, w, focal, cam['near'])
    view_dirs = rays_d / np.linalg.norm(rays_d, ord=2, axis=-1, keepdims=True)
    ones = np.ones((num_rays, 1), np.float32)
    batch = dict(rays_o=rays_o, rays_d=rays_d, view_dirs=view_dirs.astype(np.float32), rays_o_ndc=o_ndc,
                 rays_d_ndc=d_ndc, near=cam['near'] * ones, far=np.float32(cam['far']) * ones,
                 near_ndc=0 * ones, far_ndc=ones)
    out = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in batch.items()}
    out['iter_num'] = 0
    out['num_frames'] = 3
    return out


def densify_state(state: Dict[str, torch.Tensor], scale: float = 30.0, shift: float = 5.0) -> Dict[str, torch.Tensor]:
    """random-init fields are almost empty (acc ~ 4e-3), which makes
    depth = sum(w z)/(acc+1e-6) ill-conditioned.  Scaling/shifting channel 0 of every
    ``pts_output_linear`` gives an opaque field (acc ~ 1) for conditioned parity checks."""
    out = copy.copy(state)
    for name, t in state.items():
        if name.endswith('pts_output_linear.weight'):
            t = t.clone()
            t[0] *= scale
            out[name] = t
        elif name.endswith('pts_output_linear.bias'):
            t = t.clone()
            t[0] += shift
            out[name] = t
    return out


def deterministic_state(shapes: Dict[str, tuple], seed: int) -> Dict[str, torch.Tensor]:
    """Host-independent stand-in for nn.Linear's default init: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) drawn from
    numpy's PCG64 (stable across torch builds), in sorted-name order.  ``shapes``: state_dict name -> shape."""
    rng = np.random.Generator(np.random.PCG64(seed))
    state = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        fan_in = shape[-1] if name.endswith('weight') else tuple(shapes[name[:-4] + 'weight'])[-1]
        bound = 1.0 / np.sqrt(fan_in)
        state[name] = torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))
    return state


def trained_scale_state(state: Dict[str, torch.Tensor], gain: float = 6.0 ** 0.5) -> Dict[str, torch.Tensor]:
    """Weights of trained magnitude for parity fixtures: nn.Linear's default init U(-1/sqrt(fan_in), 1/sqrt(fan_in)) has a third of
    the variance that keeps a ReLU trunk's signal alive, so random-init outputs are nearly constant (rgb within 0.02 of 0.5).
    Every weight matrix times sqrt(6) is He-scaled: rgb then spans (0, 1) and sigma is O(1-10).  Biases are kept."""
    return {k: (v * gain if k.endswith('weight') else v.clone()) for k, v in state.items()}
