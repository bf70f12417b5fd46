"""Generates tests/golden/render_visibility.npz by executing the UNMODIFIED reference model with the secondary-view
visibility head switched on (`predict_visibility=True` on the coarse and fine MLPs):
four-row view head, the view branch once more per other view with the directions of `compute_other_view_dirs`, and the
`visibility2` compositing.  Case A hands `rays_o2` in (Tester contract), case B lets the model derive it from
`common_data['poses']`, `pixel_id` and `num_frames` (training contract).  Build container only:
    python oracle/make_golden_visibility.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from make_golden import GOLD, Replay, checksum, draw_randoms, full_state, ref_mod, save   # noqa: E402  (imports the real reference)
from oracle import nerf_oracle as orc                                                      # noqa: E402
from simplenerf_b200 import synthetic                                                      # noqa: E402

GRAD_KEYS = ('rgb_coarse', 'rgb_fine', 'depth_coarse', 'depth_fine', 'visibility2_coarse', 'visibility2_fine')


def visibility_configs(ndc):
    configs = synthetic.make_configs('vanilla', ndc=ndc)
    for k in ('coarse_mlp', 'fine_mlp'):
        configs['model'][k]['predict_visibility'] = True
    return configs


def main():
    arrays = {}
    for tag, (ndc, seed, given) in {'a': (True, 41, True), 'b': (False, 42, False)}.items():
        n = 10
        configs = visibility_configs(ndc)
        state = full_state(configs, seed, dense=True)
        batch = synthetic.make_ray_batch('llff', n, seed)
        g = torch.Generator().manual_seed(seed)
        poses = torch.eye(4)[None].repeat(3, 1, 1)
        poses[:, :3, 3] = torch.rand((3, 3), generator=g) - 0.5
        if given:
            batch['rays_o2'] = torch.rand((n, 2, 3), generator=g) - 0.5
        else:
            batch['pixel_id'] = torch.stack([torch.randint(0, 3, (n,), generator=g), torch.randint(0, 1008, (n,), generator=g),
                                             torch.randint(0, 756, (n,), generator=g)], -1).int()
            batch['common_data'] = {'poses': poses[None]}          # leading replica dimension, stripped by forward (:69-73)
        table = draw_randoms(configs, n, seed + 1)
        model = ref_mod.SimpleNeRF(configs, None)
        model.load_state_dict(state)
        arrays.update({f'{tag}_in_{k}': v for k, v in batch.items() if isinstance(v, torch.Tensor)})
        if not given:
            arrays[f'{tag}_in_poses'] = poses
        arrays.update({f'{tag}_rnd_{k}': v for k, v in table.items()})
        arrays[f'{tag}_checksum'] = checksum(state)
        arrays[f'{tag}_meta'] = np.array([seed, n, int(ndc), int(given)])
        model.eval()
        with torch.no_grad():
            res = model(batch, retraw=True, sec_views_vis=True)
            for k, v in res.items():
                arrays[f'{tag}_eval__{k}'] = v
            plain = model(batch)                                   # Tester contract without the head's outputs
            arrays[f'{tag}_evalplain_keys'] = np.array([len(plain), int(any('visibility2' in k for k in plain))])
        model.train()
        order = [s for s in orc.model_slots(configs) if 'fine' not in s] + [s for s in orc.model_slots(configs) if 'fine' in s]
        noise_stream = torch.cat([table[f'noise_{s}'].flatten() for s in order])
        with Replay([table['t_rand'], table['u']], noise_stream):
            res = model(batch)
        rng = np.random.Generator(np.random.PCG64(seed + 2))
        loss = 0
        for k in GRAD_KEYS:
            cot = torch.from_numpy(rng.standard_normal(tuple(res[k].shape), dtype=np.float32))
            arrays[f'{tag}_cot__{k}'] = cot
            loss = loss + (res[k] * cot).sum()
        loss.backward()
        for k, v in res.items():
            arrays[f'{tag}_train__{k}'] = v
        pick = np.random.Generator(np.random.PCG64(seed + 3))
        for pname, prm in model.named_parameters():
            gflat = prm.grad.flatten()
            idx = pick.integers(0, gflat.numel(), size=min(32, gflat.numel()))
            arrays[f'{tag}_gidx__{pname}'] = idx
            arrays[f'{tag}_gval__{pname}'] = gflat[idx]
            arrays[f'{tag}_gnorm__{pname}'] = np.array([float(gflat.double().norm()), float(gflat.double().sum())])
        print(tag, sorted(k for k in res if 'visib' in k), {k: tuple(res[k].shape) for k in res if 'visibility2' in k})
    for k in list(arrays):       # the bulky per-sample tensors nobody consumes
        if '__' in k and ('alpha' in k or 'raw_rgb' in k or 'raw_sigma' in k):
            del arrays[k]
    save('render_visibility.npz', **arrays)


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    main()
