"""Generate tests/golden/rays.npz by executing the UNMODIFIED reference ray construction and output post-processing
(src/data_preprocessors/DataPreprocessor01.py: get_rays :351-368, get_ndc_rays :371-389, get_view_dirs :392-394,
post_process_image :1106-1109, post_process_depth :1112-1114).  Build container only (needs /root/reference):

    python oracle/make_golden_rays.py

The reference module imports plotting / image-io packages that this image lacks; they are stubbed (none of them is
touched by the functions called here).  A deterministic subset of pixels of each camera is stored."""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/src')
for name in ('skimage', 'skimage.io', 'skimage.transform', 'simplejson', 'deepdiff', 'matplotlib', 'matplotlib.pyplot', 'skvideo',
             'skvideo.io', 'pandas'):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:       # noqa: BLE001
            sys.modules[name] = types.ModuleType(name)
if not hasattr(sys.modules['deepdiff'], 'DeepDiff'):
    sys.modules['deepdiff'].DeepDiff = object

from data_preprocessors import DataPreprocessor01 as ref_mod      # noqa: E402  (the real reference)
from simplenerf_b200 import synthetic                               # noqa: E402

Ref = ref_mod.DataPreprocessor
fake_self = types.SimpleNamespace(mip_nerf_used=False)


def pose_of(seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.normal(size=(3, 3))
    q, _ = np.linalg.qr(a)
    if np.linalg.det(q) < 0:
        q[:, 0] *= -1
    # mostly forward-looking camera: blend with the identity so that rays_d[..., 2] stays negative (NDC assumes it)
    r = 0.85 * np.eye(3) + 0.15 * q
    u, _, vt = np.linalg.svd(r)
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = (u @ vt).astype(np.float32)
    pose[:3, 3] = rng.normal(size=3).astype(np.float32) * 0.1
    return pose


out = {}
for cam_name, seed in (('llff', 5), ('re10k', 6)):
    cam = synthetic.CAMERAS[cam_name]
    h, w = cam['resolution']
    intrinsic = np.array([[cam['focal'], 0, cam['centre'][0]], [0, cam['focal'] * 1.01, cam['centre'][1]], [0, 0, 1]], dtype=np.float32)
    pose = pose_of(seed)
    rays_o, rays_d = Ref.get_rays(fake_self, (h, w), intrinsic, pose)
    view_dirs = Ref.get_view_dirs(rays_d)
    o_ndc, d_ndc = Ref.get_ndc_rays(rays_o, rays_d, (h, w), intrinsic, cam['near'])
    pick = np.arange(0, h * w, 997)
    out[f'{cam_name}_intrinsic'] = intrinsic
    out[f'{cam_name}_pose'] = pose
    out[f'{cam_name}_pick'] = pick
    for key, arr in (('rays_o', rays_o), ('rays_d', rays_d), ('view_dirs', view_dirs), ('rays_o_ndc', o_ndc), ('rays_d_ndc', d_ndc)):
        out[f'{cam_name}_{key}'] = np.ascontiguousarray(np.reshape(arr, (-1, 3))[pick]).astype(np.float32)
        out[f'{cam_name}_{key}_dtype'] = np.array(str(arr.dtype))
rng = np.random.Generator(np.random.PCG64(11))
rgb = np.concatenate([rng.uniform(-0.2, 1.2, size=(4000, 3)), (np.arange(0, 256).reshape(-1, 1) + 0.5) / 255 * np.ones((1, 3)),
                      np.arange(0, 256).reshape(-1, 1) / 255 * np.ones((1, 3))]).astype(np.float32)
depth = rng.normal(size=5000).astype(np.float32) * 3
out['post_rgb'] = rgb
out['post_image'] = Ref.post_process_image(rgb)
out['post_depth_in'] = depth
out['post_depth'] = Ref.post_process_depth(depth)
path = os.path.join(ROOT, 'tests', 'golden', 'rays.npz')
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), {k: str(out[k]) for k in out if k.endswith('_dtype')})
