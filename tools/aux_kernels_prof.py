"""One launch of every non-MLP kernel of the path (samplers, compositing, losses, batch gather) for an ncu capture:
   ncu --set full -k "regex:sample_|composite_|ray_losses|reproj_|gather_rows" python tools/aux_kernels_prof.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from simplenerf_b200 import ops
from simplenerf_b200.batching import gather_rows
from simplenerf_b200.loss_functions import ray_losses, reprojection_losses

DEV = 'cuda:0'
n = 1 << int(os.environ.get('LOGN', '20'))
g = torch.Generator(device=DEV).manual_seed(1)
r = lambda *s: torch.rand(s, device=DEV, generator=g)      # noqa: E731
for s in (64, 192):
    sigma, rgb = torch.relu(3 * torch.randn((n, s), device=DEV, generator=g)), r(n, s, 3)
    z = torch.sort(r(n, s), -1)[0].contiguous()
    o, d = torch.randn((n, 3), device=DEV, generator=g), torch.nn.functional.normalize(torch.randn((n, 3), device=DEV, generator=g), dim=-1) * 2
    d[:, 2] = -d[:, 2].abs() - 0.1
    ops.composite_forward(sigma, rgb, z, o, d, d, True, False, per_sample=('weights',))
    ops.composite_backward(sigma, rgb, z, o, d, d, True, False, {'rgb': r(n, 3), 'depth': r(n)})
    if s == 64:
        w = r(n, 64)
        ops.sample_fine(z, w, r(n, 128))
        ops.sample_fine(z, w, torch.linspace(0, 1, 128).to(DEV))
        ops.sample_coarse(torch.zeros(n, device=DEV), torch.ones(n, device=DEV), torch.linspace(0, 1, 64).to(DEV), r(n, 64))
    del sigma, rgb, z
# the training step's own sizes for the loss / batch kernels
m = 4096
preds = [r(m, 3).requires_grad_() for _ in range(4)] + [r(m).requires_grad_() for _ in range(4)]
mask = r(m) < 0.75
vals = ray_losses(preds, [r(m, 3)] * 4 + [r(m)] * 4, [mask] * 4 + [~mask] * 4, [1.0] * 4 + [0.1] * 4)
vals[-1].backward()
import numpy as np
with np.load(os.path.join(ROOT, 'tests', 'golden', 'losses.npz')) as _f:      # inputs only: a fixture file, no oracle code
    f = {k: torch.from_numpy(_f[k]) for k in _f.files if k.startswith('r_')}
inp = {k[5:]: v.to(DEV) for k, v in f.items() if k.startswith('r_in_')}
cd = {k[9:]: v.to(DEV) for k, v in f.items() if k.startswith('r_common_')}
out = {k[6:]: v.to(DEV).requires_grad_() for k, v in f.items() if k.startswith('r_out_')}
v, _ = reprojection_losses(out['depth_coarse'], [out['points_augmentation_depth_coarse'], out['views_augmentation_depth_coarse'], out['depth_fine']],
                           [0.1] * 3, inp['rays_o'], inp['rays_d'], inp['pixel_id'], inp['indices_mask_nerf'], cd['images'], cd['poses'], cd['intrinsics'])
v[-1].backward()
src = [r(3 * 756 * 1008, 3) for _ in range(4)]
gather_rows([(t, torch.empty((m, 3), device=DEV), None) for t in src], torch.randint(0, 3 * 756 * 1008, (m,), device=DEV))
torch.cuda.synchronize()
print('ok')
