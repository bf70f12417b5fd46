"""Cost of the visibility head on the tensor path: a vanilla (coarse + fine) training step of 4096 rays with and without
predict_visibility (3 views: two other views per point), CUDA events.  python tools/vis_prof.py [steps]"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from simplenerf_b200 import synthetic  # noqa: E402
from simplenerf_b200.models import get_model  # noqa: E402


def run(vis: bool, steps: int, n: int = 4096):
    dev = torch.device('cuda', 0)
    configs = synthetic.make_configs('vanilla')
    if vis:
        for k in ('coarse_mlp', 'fine_mlp'):
            configs['model'][k]['predict_visibility'] = True
    model = get_model(configs, None)
    model.load_state_dict(bench.make_state(model))
    model = model.to(dev).train()
    batch = synthetic.make_ray_batch('llff', n, 1021)
    g = torch.Generator().manual_seed(3)
    batch['rays_o2'] = torch.rand((n, 2, 3), generator=g) - .5
    target = torch.rand((n, 3), generator=g).to(dev)
    batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}

    def step():
        model.zero_grad(set_to_none=True)
        out = model(batch)
        loss = ((out['rgb_coarse'] - target) ** 2).mean() + ((out['rgb_fine'] - target) ** 2).mean()
        if vis:      # VisibilityLoss01 / VisibilityPriorLoss01 shaped consumers of the head
            for lvl in ('coarse', 'fine'):
                loss = loss + (out[f'raw_visibility_{lvl}'][..., 0] - out[f'visibility_{lvl}'].detach()).abs().mean()
                loss = loss + 0.01 * (1 - out[f'visibility2_{lvl}']).sum(1).mean()
        loss.backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


if __name__ == '__main__':
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    a = run(False, steps)
    b = run(True, steps)
    print(f'vanilla step, 4096 rays x (64 + 192) points: {a:.3f} ms; with the visibility head (2 other views): {b:.3f} ms (+{b - a:.3f} ms)')
