"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): one tiny training step of the 4-MLP model, one fused evaluation
(NDC and metric rays), the 256-sample compositing backward.  Sizes keep a sanitizer run in the minutes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from simplenerf_b200 import ops, synthetic
from simplenerf_b200.models import get_model

DEV = 'cuda:0'


def build(kind, ndc=True):
    cfg = synthetic.make_configs(kind, ndc=ndc)
    model = get_model(cfg, None)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synthetic.densify_state(synthetic.deterministic_state(shapes, 1)))
    return model.to(DEV)


def dev(batch):
    return {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


n = 96
model = build('simplenerf').train()
out = model(dev(synthetic.make_ray_batch('llff', n, 3)))
sum(out[k].sum() for k in ('rgb_coarse', 'rgb_fine', 'depth_coarse', 'points_augmentation_rgb_coarse', 'views_augmentation_depth_coarse')).backward()
torch.cuda.synchronize()
print('training step ok')
for cam, ndc in (('llff', True), ('re10k', False)):
    with torch.no_grad():
        o = build('vanilla', ndc).eval()(dev(synthetic.make_ray_batch(cam, n + 5, 4)))
    torch.cuda.synchronize()
    assert all(bool(torch.isfinite(v).all()) for v in o.values())
    print('fused evaluation ok', cam)
g = torch.Generator().manual_seed(0)
m, s = 37, 256
b = synthetic.make_ray_batch('llff', m, 1)
sigma, rgb = torch.rand((m, s), generator=g) * 3, torch.rand((m, s, 3), generator=g)
z = torch.sort(torch.rand((m, s), generator=g), -1)[0]
grads = {'rgb': torch.randn((m, 3), generator=g).to(DEV), 'depth': torch.randn(m, generator=g).to(DEV)}
ops.composite_backward(sigma.to(DEV), rgb.to(DEV), z.to(DEV).contiguous(), b['rays_o'].to(DEV), b['rays_d'].to(DEV), b['rays_d_ndc'].to(DEV), True, False, grads)
torch.cuda.synchronize()
print('compositing backward (256 samples) ok')
