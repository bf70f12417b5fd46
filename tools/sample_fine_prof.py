"""One launch of the hierarchical resampler per variant of the uniforms (random / sorted / linspace row) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from simplenerf_b200 import ops

n = 1 << int(os.environ.get('LOGN', '20'))
g = torch.Generator(device='cuda:0').manual_seed(1)
z = torch.sort(torch.rand((n, 64), device='cuda:0', generator=g), -1)[0].contiguous()
w = torch.rand((n, 64), device='cuda:0', generator=g)
u = torch.rand((n, 128), device='cuda:0', generator=g)
lin = torch.linspace(0, 1, 128).to('cuda:0')
for _ in range(2):
    ops.sample_fine(z, w, u)
    ops.sample_fine(z, w, lin)
torch.cuda.synchronize()
