"""Stress check of the asynchronous protocols: repeats the forward / backward of one MLP on fixed inputs and requires
bit-identical sigma / rgb every time and gradients that agree to fp32-atomics noise."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from simplenerf_b200 import ops, synthetic
from simplenerf_b200._lib import FLAG_SAVE_FOR_BWD
from simplenerf_b200.models.FusedSimpleNeRF01 import MlpBlock
DEV = 'cuda:0'
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
model_cfg = synthetic.make_configs('simplenerf')['model']
for name, cfg, n_rays, s in (('main', model_cfg['coarse_mlp'], 4096, 64), ('pts-aug', model_cfg['points_augmentation']['coarse_mlp'], 3000, 64),
                             ('views-aug', model_cfg['views_augmentation']['coarse_mlp'], 4096, 64), ('fine', model_cfg['fine_mlp'], 1111, 192)):
    torch.manual_seed(0)
    block = MlpBlock(cfg).to(DEV)
    table = [None if p is None else p.detach() for p in block.param_table()]
    packed = block.packed(table)
    b = synthetic.make_ray_batch('llff', n_rays, 3)
    o, d, vd = b['rays_o_ndc'].to(DEV), b['rays_d_ndc'].to(DEV), b['view_dirs'].to(DEV)
    z = torch.sort(torch.rand(n_rays, s, device=DEV), -1)[0].contiguous()
    noise = torch.randn(n_rays * s, device=DEV)
    flags = FLAG_SAVE_FOR_BWD
    ws = torch.empty(ops.mlp_workspace_bytes(block.desc, n_rays, s, flags), dtype=torch.uint8, device=DEV)
    ds = torch.randn(n_rays, s, device=DEV); dr = torch.randn(n_rays, s, 3, device=DEV)
    ref = None
    worst = 0.0
    for it in range(reps):
        sigma, rgb = ops.mlp_forward(block.desc, table, packed, o, d, vd, z, noise, ws, flags)
        grads = [None if p is None else torch.zeros_like(p) for p in table]
        ops.mlp_backward(block.desc, table, packed, o, d, vd if block.view_degree else None, z, sigma, rgb, ds, dr, grads, ws, flags)
        cur = (sigma.clone(), rgb.clone(), [None if g is None else g.clone() for g in grads])
        assert bool(torch.isfinite(sigma).all()) and bool(torch.isfinite(rgb).all())
        if ref is None:
            ref = cur
        else:
            assert torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1]), f'{name}: forward not reproducible at repetition {it}'
            for ga, gb in zip(cur[2], ref[2]):
                if ga is not None:
                    rel = float((ga - gb).norm() / (gb.norm() + 1e-20))
                    worst = max(worst, rel)
                    assert rel < 1e-4, f'{name}: gradient differs between repetitions ({rel})'
    print(f'{name}: {reps} repetitions identical forward, gradient repeatability {worst:.2e}', flush=True)
print('STRESS OK')
