import os, sys
sys.path.insert(0, '.')
import torch
from simplenerf_b200 import ops
DEV='cuda:0'; n=1<<19; s=256
g = torch.Generator(device=DEV).manual_seed(1)
r = lambda *sh: torch.rand(sh, device=DEV, generator=g)
sigma, rgb = torch.relu(3 * torch.randn((n, s), device=DEV, generator=g)), r(n, s, 3)
z = torch.sort(r(n, s), -1)[0].contiguous()
o, d = torch.randn((n, 3), device=DEV, generator=g), torch.nn.functional.normalize(torch.randn((n, 3), device=DEV, generator=g), dim=-1) * 2
d[:, 2] = -d[:, 2].abs() - 0.1
for _ in range(2):
    ops.composite_backward(sigma, rgb, z, o, d, d, True, False, {'rgb': r(n, 3), 'depth': r(n)})
torch.cuda.synchronize()
