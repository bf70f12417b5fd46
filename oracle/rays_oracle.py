"""TEST INFRASTRUCTURE (checker only; never on the product path).

numpy restatement of the reference's per-frame ray construction and output post-processing, pinned against
tests/golden/rays.npz (generated from the unmodified reference by oracle/make_golden_rays.py):
  get_rays            src/data_preprocessors/DataPreprocessor01.py:351-368
  get_ndc_rays        :371-389
  get_view_dirs       :392-394
  post_process_image  :1106-1109      post_process_depth :1112-1114
All arithmetic is numpy fp32 under NEP 50 (python scalars are weak), which is what the reference executes with the numpy 2
of this image."""
from __future__ import annotations

import numpy as np


def get_rays(resolution, intrinsic: np.ndarray, pose: np.ndarray):
    h, w = resolution
    x, y = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32), indexing='xy')       # :353-356
    points_homo = np.stack([x, y, np.ones_like(x)], axis=2)                                                 # :360-361
    dirs = (np.linalg.inv(intrinsic)[None, None] @ points_homo[:, :, :, None])[:, :, :, 0]                  # :362
    dirs[:, :, 1:] *= -1                                                                                    # :363
    rays_d = np.sum(dirs[..., np.newaxis, :] * pose[:3, :3], -1)                                            # :365
    rays_o = np.broadcast_to(pose[:3, -1], np.shape(rays_d))                                                # :367
    return rays_o, rays_d


def get_ndc_rays(rays_o, rays_d, resolution, intrinsic, near):
    h, w = resolution
    fx, fy = intrinsic[0, 0], intrinsic[1, 1]
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]                                                           # :375
    rays_o = rays_o + t[..., None] * rays_d                                                                 # :376
    o0 = -1. / (w / (2. * fx)) * rays_o[..., 0] / rays_o[..., 2]                                            # :379
    o1 = -1. / (h / (2. * fy)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (w / (2. * fx)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])        # :383
    d1 = -1. / (h / (2. * fy)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2. * near / rays_o[..., 2]
    return np.stack([o0, o1, o2], -1), np.stack([d0, d1, d2], -1)


def get_view_dirs(rays_d):
    return rays_d / np.linalg.norm(rays_d, ord=2, axis=-1, keepdims=True)                                   # :393


def post_process_image(rgb):
    return np.round(np.clip(rgb, a_min=0, a_max=1) * 255).astype('uint8')                                   # :1107-1108


def post_process_depth(depth):
    return np.clip(depth, a_min=0, a_max=np.inf).astype('float32')                                          # :1113


def frame_rays(resolution, intrinsic, pose, near, ndc=True):
    """The ray part of create_test_data (:807-866), flattened to [H*W, 3] fp32."""
    intrinsic = np.asarray(intrinsic, dtype=np.float32)
    pose = np.asarray(pose, dtype=np.float32)
    rays_o, rays_d = get_rays(resolution, intrinsic, pose)
    out = {'rays_o': rays_o, 'rays_d': rays_d, 'view_dirs': get_view_dirs(rays_d)}
    if ndc:
        out['rays_o_ndc'], out['rays_d_ndc'] = get_ndc_rays(rays_o, rays_d, resolution, intrinsic, near)
    return {k: np.ascontiguousarray(np.reshape(v, (-1, 3))).astype(np.float32) for k, v in out.items()}
