"""C5 microbench: compositing forward/backward, hierarchical resampling and stratified sampling, achieved HBM GB/s
against the algorithmic bytes of SURVEY.md section 8(d) (20S+68 / 24S+68 / 36S+48 bytes per ray, 1792 / 1280 bytes per
ray for sample_pdf+merge, 4S(+4S) bytes per ray for the stratified sampler)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from simplenerf_b200 import ops

DEV = 'cuda:0'
PEAK = 6548.2
if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')):
    PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3



def run(ss=(64, 128, 256), logns=(16, 18, 20, 22), iters=10):
    """-> list of dicts {kernel, S, rays, algorithmic_bytes, seconds, gbs, frac} (frac of the measured HBM copy bandwidth)."""
    rows = []
    for s in ss:
        for logn in logns:
            n = 1 << logn
            if n * s * 4 * 9 > 60e9:
                continue
            g = torch.Generator(device=DEV).manual_seed(1000 + logn * 10 + s // 64)
            sigma = torch.relu(3 * torch.randn((n, s), device=DEV, generator=g))
            rgb = torch.sigmoid(torch.randn((n, s, 3), device=DEV, generator=g))
            z = torch.sort(torch.rand((n, s), device=DEV, generator=g), -1)[0].contiguous()
            o = torch.randn((n, 3), device=DEV, generator=g)
            d = torch.nn.functional.normalize(torch.randn((n, 3), device=DEV, generator=g), dim=-1) * 2
            d[:, 2] = -d[:, 2].abs() - 0.1
            t = timeit(lambda: ops.composite_forward(sigma, rgb, z, o, d, d, True, False, per_sample=()), iters)
            rows.append(('composite_fwd (render contract)', s, n, (20 * s + 68) * n, t))
            t = timeit(lambda: ops.composite_forward(sigma, rgb, z, o, d, d, True, False, per_sample=('weights',)), iters)
            rows.append(('composite_fwd (+weights)', s, n, (24 * s + 68) * n, t))
            g_rgb, g_depth = torch.randn((n, 3), device=DEV), torch.randn(n, device=DEV)
            t = timeit(lambda: ops.composite_backward(sigma, rgb, z, o, d, d, True, False, {'rgb': g_rgb, 'depth': g_depth}), iters)
            rows.append(('composite_bwd', s, n, (36 * s + 48) * n, t))
            if s == 64:
                w = torch.rand((n, 64), device=DEV, generator=g)
                u = torch.rand((n, 128), device=DEV, generator=g)
                t = timeit(lambda: ops.sample_fine(z, w, u), iters)
                rows.append(('sample_pdf+merge (u supplied)', 64, n, 1792 * n, t))
                us = torch.sort(u, -1)[0].contiguous()
                t = timeit(lambda: ops.sample_fine(z, w, us), iters)
                rows.append(('sample_pdf+merge (sorted u supplied)', 64, n, 1792 * n, t))
                del us
                lin = torch.linspace(0, 1, 128).to(DEV)
                t = timeit(lambda: ops.sample_fine(z, w, lin), iters)
                rows.append(('sample_pdf+merge (linspace row)', 64, n, 1280 * n, t))
                near, far, tv = torch.zeros(n, device=DEV), torch.ones(n, device=DEV), torch.linspace(0, 1, 64).to(DEV)
                tr = torch.rand((n, 64), device=DEV, generator=g)
                t = timeit(lambda: ops.sample_coarse(near, far, tv, tr), iters)
                rows.append(('stratified sampler (t_rand supplied)', 64, n, 8 * 64 * n, t))
            del sigma, rgb, z
    return [dict(kernel=name, S=s, rays=n, algorithmic_bytes=b, seconds=t, gbs=b / t / 1e9, frac=b / t / 1e9 / PEAK)
            for name, s, n, b, t in rows]


if __name__ == '__main__':
    # optional quick mode: SCAN_S=64,256 SCAN_LOGN=20,22 restrict the sweep
    SS = tuple(int(x) for x in os.environ.get('SCAN_S', '64,128,256').split(','))
    LOGNS = tuple(int(x) for x in os.environ.get('SCAN_LOGN', '16,18,20,22').split(','))
    print(f'| kernel | S | rays | algorithmic MB | time us | GB/s | frac of measured {PEAK:.0f} GB/s |')
    print('|---|---|---|---|---|---|---|')
    for r in run(SS, LOGNS):
        print(f"| {r['kernel']} | {r['S']} | 2^{r['rays'].bit_length() - 1} | {r['algorithmic_bytes'] / 1e6:.1f} | {r['seconds'] * 1e6:.1f} | "
              f"{r['gbs']:.0f} | {r['frac']:.2f} |")
