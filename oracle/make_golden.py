"""Generate tests/golden/*.npz by executing the UNMODIFIED reference module.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

Every fixture stores the inputs, the injected random draws and the outputs of the reference's
own functions (``models.SimpleNeRF01``), so that ``tests/test_oracle_golden.py`` can pin the
oracle restatement (``oracle/nerf_oracle.py``) and the ``-m gpu`` tests can pin the CUDA path
against the real reference without the reference being present.
Weights come from ``oracle.nerf_oracle.deterministic_state`` (numpy PCG64, host independent);
a checksum of them is stored in each fixture.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/src')

from models import SimpleNeRF01 as ref_mod                      # noqa: E402  (the real reference)
from oracle import nerf_oracle as orc                             # noqa: E402
from simplenerf_b200 import synthetic                             # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
GRAD_KEYS = ('rgb_coarse', 'rgb_fine', 'depth_coarse', 'depth_fine',
             'points_augmentation_rgb_coarse', 'points_augmentation_depth_coarse',
             'views_augmentation_rgb_coarse', 'views_augmentation_depth_coarse',
             # fine-level augmentation models (:234-263): present only in the 'simplenerf_fineaug' fixture; appended, so the
             # cotangents of the older fixtures are drawn exactly as before
             'points_augmentation_rgb_fine', 'points_augmentation_depth_fine',
             'views_augmentation_rgb_fine', 'views_augmentation_depth_fine')


def full_state(configs, seed, dense=False):
    shapes = {}
    for slot, mlp_cfg in orc.model_slots(configs).items():
        for k, v in orc.MlpSpec(mlp_cfg).param_shapes().items():
            shapes[f'{slot}.{k}'] = v
    state = orc.deterministic_state(shapes, seed)
    return synthetic.densify_state(state) if dense else state


def checksum(state):
    return np.array([sum(float(v.double().sum()) for v in state.values()),
                     sum(float(v.double().abs().sum()) for v in state.values())])


class Replay:
    """Hands pre-drawn tensors to the reference's torch.rand / torch.randn calls in call order."""

    def __init__(self, rand_queue, randn_stream):
        self.rand_queue = list(rand_queue)
        self.randn_stream = randn_stream
        self.pos = 0

    def rand(self, shape, *a, **k):
        t = self.rand_queue.pop(0)
        assert list(t.shape) == list(shape), (t.shape, shape)
        return t.clone()

    def randn(self, shape, *a, **k):
        n = int(np.prod(list(shape)))
        t = self.randn_stream[self.pos:self.pos + n].reshape(list(shape))
        self.pos += n
        return t.clone()

    def __enter__(self):
        self._rand, self._randn = torch.rand, torch.randn
        torch.rand, torch.randn = self.rand, self.randn
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randn = self._rand, self._randn


def draw_randoms(configs, n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    s_c = configs['model']['coarse_mlp']['num_samples']
    n_f = configs['model']['fine_mlp']['num_samples']
    table = {'t_rand': rng.random((n, s_c), dtype=np.float32), 'u': rng.random((n, n_f), dtype=np.float32)}
    for slot in orc.model_slots(configs):
        s = s_c + n_f if 'fine' in slot else s_c
        table[f'noise_{slot}'] = rng.standard_normal((n * s, 1), dtype=np.float32)
    return {k: torch.from_numpy(v) for k, v in table.items()}


def save(name, **arrays):
    flat = {}
    for k, v in arrays.items():
        flat[k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    path = os.path.join(GOLD, name)
    np.savez_compressed(path, **flat)
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(flat)} arrays')


# ------------------------------------------------------------------------------------------------
def golden_ops():
    rng = np.random.Generator(np.random.PCG64(11))
    out = {}
    # a12 sample_pdf: random, deterministic, and degenerate weights (all-zero row, one-hot row, tiny row)
    n, b = 12, 63
    z = np.sort(rng.random((n, 64), dtype=np.float32), axis=-1)
    bins = torch.from_numpy(.5 * (z[:, 1:] + z[:, :-1]))
    w = rng.random((n, b - 1), dtype=np.float32) ** 4
    w[0] = 0
    w[1] = 0
    w[1, 17] = 1
    w[2] = 1e-9
    w[3, :40] = 0
    w = torch.from_numpy(w)
    u = torch.from_numpy(rng.random((n, 128), dtype=np.float32))
    u[4, :4] = torch.tensor([0., 1e-8, 0.99999994, 0.5])
    with Replay([u], None):
        out['pdf_rand'] = ref_mod.SimpleNeRF.sample_pdf(bins, w, 128, det=False)
    out['pdf_det'] = ref_mod.SimpleNeRF.sample_pdf(bins, w, 128, det=True)
    out.update(pdf_bins=bins, pdf_weights=w, pdf_u=u)
    # a10
    o = torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32))
    d = torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32))
    d[:, 2] = -d[:, 2].abs() - .2
    zn = torch.from_numpy(z.copy())
    zn[:, -1] = 1.
    out.update(ndc_o=o, ndc_d=d, ndc_z=zn, ndc_depth=ref_mod.SimpleNeRF.convert_depth_from_ndc(zn, o, d))
    # a6 positional encoding (degree 10 and 4)
    x = torch.from_numpy(rng.uniform(-1.5, 1.5, size=(40, 3)).astype(np.float32))
    for deg in (10, 4, 3):
        fn, dim = ref_mod.MLP.get_positional_encoder(deg)
        out[f'pe{deg}'] = fn(x)
    out['pe_x'] = x
    save('ops.npz', **out)


def golden_mlp():
    configs = synthetic.make_configs('simplenerf')
    rng = np.random.Generator(np.random.PCG64(5))
    p = 96
    pts = torch.from_numpy(rng.uniform(-1.2, 1.2, size=(p, 3)).astype(np.float32))
    vd = rng.normal(size=(p, 3)).astype(np.float32)
    vd = torch.from_numpy(vd / np.linalg.norm(vd, axis=-1, keepdims=True))
    noise = torch.from_numpy(rng.standard_normal((p, 1), dtype=np.float32))
    out = dict(pts=pts, view_dirs=vd, noise=noise)
    for slot, mlp_cfg in orc.model_slots(configs).items():
        spec = orc.MlpSpec(mlp_cfg)
        state = orc.deterministic_state(spec.param_shapes(), 100 + len(slot))
        mlp = ref_mod.MLP(configs, mlp_cfg)
        mlp.load_state_dict(state)
        for training in (False, True):
            mlp.train(training)
            with Replay([], noise.flatten()):
                res = mlp({'pts': pts, 'view_dirs': vd})
            tag = f"{slot}_{'train' if training else 'eval'}"
            out[f'{tag}_sigma'] = res['sigma']
            out[f'{tag}_rgb'] = res['rgb']
        out[f'{slot}_checksum'] = checksum(state)
    save('mlp.npz', **out)


def golden_render(name, kind, ndc, camera, n, dense, seed):
    configs = synthetic.make_configs(kind, ndc=ndc)
    state = full_state(configs, seed, dense)
    batch = synthetic.make_ray_batch(camera, n, seed)
    table = draw_randoms(configs, n, seed + 1)
    model = ref_mod.SimpleNeRF(configs, None)
    model.load_state_dict(state)
    arrays = {f'in_{k}': v for k, v in batch.items() if isinstance(v, torch.Tensor)}
    arrays.update({f'rnd_{k}': v for k, v in table.items()})
    arrays['checksum'] = checksum(state)
    arrays['meta'] = np.array([seed, n, int(dense), int(ndc)])

    # eval (Tester contract: retraw False) and eval with retraw (validation contract)
    model.eval()
    with torch.no_grad():
        for retraw in (False, True):
            res = model(batch, retraw=retraw)
            for k, v in res.items():
                arrays[f"eval{'_raw' if retraw else ''}__{k}"] = v

    # train: forward + backward through the 8 grad-carrying outputs with fixed cotangents
    model.train()
    slots = list(orc.model_slots(configs))
    s_c, n_f = 64, 128
    noise_stream = torch.cat([table[f'noise_{s}'].flatten() for s in slots])   # coarse.. then fine (ctor order)
    order = [s for s in slots if 'fine' not in s] + [s for s in slots if 'fine' in s]
    noise_stream = torch.cat([table[f'noise_{s}'].flatten() for s in order])
    with Replay([table['t_rand'], table['u']], noise_stream):
        res = model(batch)
    rng = np.random.Generator(np.random.PCG64(seed + 2))
    loss = 0
    for k in GRAD_KEYS:
        if k in res:
            cot = torch.from_numpy(rng.standard_normal(tuple(res[k].shape), dtype=np.float32))
            arrays[f'cot__{k}'] = cot
            loss = loss + (res[k] * cot).sum()
    loss.backward()
    for k, v in res.items():
        arrays[f'train__{k}'] = v
    pick = np.random.Generator(np.random.PCG64(seed + 3))
    for pname, prm in model.named_parameters():
        g = prm.grad.flatten()
        idx = pick.integers(0, g.numel(), size=min(48, g.numel()))
        arrays[f'gidx__{pname}'] = idx
        arrays[f'gval__{pname}'] = g[idx]
        arrays[f'gnorm__{pname}'] = np.array([float(g.double().norm()), float(g.double().sum())])
    # drop the bulky per-sample tensors nobody consumes (kept: weights, z_vals for teacher forcing)
    for k in list(arrays):
        if ('alpha' in k or 'raw_rgb_view' in k) and '__' in k:
            del arrays[k]
    save(name, **arrays)


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    golden_ops()
    golden_mlp()
    golden_render('render_llff_simplenerf.npz', 'simplenerf', True, 'llff', 12, False, 1021)
    golden_render('render_llff_simplenerf_dense.npz', 'simplenerf', True, 'llff', 12, True, 1022)
    golden_render('render_re10k_vanilla_dense.npz', 'vanilla', True, 're10k', 12, True, 21)
    golden_render('render_nondc_vanilla_dense.npz', 'vanilla', False, 'llff', 12, True, 33)
    golden_render('render_llff_fineaug_dense.npz', 'simplenerf_fineaug', True, 'llff', 12, True, 1023)
