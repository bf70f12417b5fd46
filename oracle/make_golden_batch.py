"""Generates tests/golden/batch.npz by calling the UNMODIFIED reference methods
DataPreprocessor.load_nerf_cached_batch / load_sparse_depth_cached_batch (src/data_preprocessors/DataPreprocessor01.py
:572-620, :655-700) on a stand-in `self` that carries only the attributes those methods read (preprocessed_data_dict,
device, ndc).  Modules the reference imports but these methods never touch (skimage) are stubbed.  Build container only:
    python oracle/make_golden_batch.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for name in ('skimage', 'skimage.io', 'skimage.transform'):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, '/root/reference/src')
from data_preprocessors.DataPreprocessor01 import DataPreprocessor  # noqa: E402


def tables(m, seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(s, generator=g)      # noqa: E731
    nerf = dict(rays_o=r(m, 3), rays_d=r(m, 3), view_dirs=r(m, 3), pixel_id=torch.randint(0, 1000, (m, 3), generator=g, dtype=torch.int32),
                target_rgb=r(m, 3), near_array=r(m, 1), far_array=r(m, 1), rays_o_ndc=r(m, 3), rays_d_ndc=r(m, 3),
                near_array_ndc=r(m, 1), far_array_ndc=r(m, 1))
    sd = dict(depths=r(m, 1), reprojection_errors=r(m, 1), depths_ndc=r(m, 1))
    return nerf, sd


def main():
    store = {}
    for tag, (m, n_nerf, n_sd, seed) in {'a': (800, 300, 100, 1), 'b': (64, 7, 0, 2)}.items():
        nerf, sd = tables(m, seed)
        g = torch.Generator().manual_seed(seed + 50)
        indices = torch.randint(0, m, (n_nerf + n_sd,), generator=g)
        ids = torch.cat([torch.ones(n_nerf), 2 * torch.ones(n_sd)])
        fake = types.SimpleNamespace(preprocessed_data_dict=dict(nerf_data=nerf, sparse_depth_data=sd, frame_nums=np.arange(3)),
                                     device='cpu', ndc=True)
        indices_dict = dict(indices=indices, indices_mask_nerf=ids == 1)
        if n_sd:
            indices_dict['indices_mask_sparse_depth'] = ids == 2
        batch = DataPreprocessor.load_nerf_cached_batch(fake, 5, indices_dict)
        batch.update(DataPreprocessor.load_sparse_depth_cached_batch(fake, indices_dict, batch))
        for k, v in nerf.items():
            store[f'{tag}_nerf_{k}'] = v.numpy()
        for k, v in sd.items():
            store[f'{tag}_sd_{k}'] = v.numpy()
        for k, v in indices_dict.items():
            store[f'{tag}_idx_{k}'] = v.numpy()
        for k, v in batch.items():
            if isinstance(v, torch.Tensor):
                store[f'{tag}_batch_{k}'] = v.numpy()
        print(tag, sorted(k for k, v in batch.items() if isinstance(v, torch.Tensor)))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'batch.npz'), **store)


if __name__ == '__main__':
    main()
