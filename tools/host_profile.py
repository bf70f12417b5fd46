"""Where the host time of one training step goes (bench.py's step, batch resident on the device): cProfile over 20 steps,
top entries by own time.  python tools/host_profile.py [out.txt]"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from simplenerf_b200 import synthetic  # noqa: E402
from simplenerf_b200.models import get_model  # noqa: E402
from simplenerf_b200.optim import FusedAdam  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    model = get_model(synthetic.make_configs('simplenerf'), None)
    model.load_state_dict(bench.make_state(model))
    model = model.to(dev).train()
    opt = FusedAdam(model.parameters(), lr=5e-4, betas=(0.9, 0.999))
    n = bench.RAYS_PER_GPU
    batch = synthetic.make_ray_batch('llff', n, 1021)
    g = torch.Generator().manual_seed(3)
    batch['target_rgb'] = torch.rand((n, 3), generator=g)
    batch['target_depth'] = 1 + 4 * torch.rand((n,), generator=g)
    batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}

    def step():
        opt.zero_grad(set_to_none=True)
        out = model(batch)
        loss = bench.fused_training_loss(out, batch['target_rgb'], batch['target_depth'])
        loss.backward()
        opt.step()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        step()
    host = 1e3 * (time.perf_counter() - t0) / 20
    torch.cuda.synchronize()
    wall = 1e3 * (time.perf_counter() - t0) / 20
    prof = cProfile.Profile()
    prof.enable()
    for _ in range(20):
        step()
    prof.disable()
    torch.cuda.synchronize()
    buf = io.StringIO()
    buf.write(f'host enqueue {host:.3f} ms per step, wall {wall:.3f} ms per step (20 steps)\n')
    pstats.Stats(prof, stream=buf).sort_stats('tottime').print_stats(45)
    pstats.Stats(prof, stream=buf).sort_stats('cumulative').print_stats(45)
    text = buf.getvalue()
    if len(sys.argv) > 1:
        with open(sys.argv[1], 'w') as f:
            f.write(text)
    print(text[:6000])


if __name__ == '__main__':
    main()
