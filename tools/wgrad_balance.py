"""Per-job time of the wgrad kernel (one job per CTA): prints, for every job, the CTAs it got and their cycle counts.
Debug tool for the static CTA allocation in tc_backward()."""
import os as _os; _os.environ['SNERF_B200_DEBUG_LIB'] = '1'   # snerfdbg_* entry points live in libsimplenerf_b200_dbg.so (build.py --debug)
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from simplenerf_b200 import ops, synthetic, _lib
from simplenerf_b200._lib import FLAG_SAVE_FOR_BWD
from simplenerf_b200.models.FusedSimpleNeRF01 import MlpBlock

DEV = 'cuda:0'
lib = _lib.load()
lib.snerfdbg_set_wgrad_trace.argtypes = [ctypes.c_void_p]
lib.snerfdbg_set_wgrad_debug.argtypes = [ctypes.c_int]
lib.snerfdbg_set_wgrad_debug(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_rays = 4096
for name, key, s in (('fine', 'fine_mlp', 192), ('views-aug', None, 64), ('pts-aug', None, 64)):
    model_cfg = synthetic.make_configs('simplenerf')['model']
    cfg = model_cfg[key] if key else (model_cfg['views_augmentation']['coarse_mlp'] if name == 'views-aug'
                                      else model_cfg['points_augmentation']['coarse_mlp'])
    block = MlpBlock(cfg).to(DEV)
    table = [None if p is None else p.detach() for p in block.param_table()]
    packed = block.packed(table)
    b = synthetic.make_ray_batch('llff', n_rays, 3)
    o, d, vd = b['rays_o_ndc'].to(DEV), b['rays_d_ndc'].to(DEV), b['view_dirs'].to(DEV)
    z = torch.sort(torch.rand(n_rays, s, device=DEV), -1)[0].contiguous()
    flags = FLAG_SAVE_FOR_BWD
    ws = torch.empty(ops.mlp_workspace_bytes(block.desc, n_rays, s, flags), dtype=torch.uint8, device=DEV)
    sigma, rgb = ops.mlp_forward(block.desc, table, packed, o, d, vd, z, None, ws, flags)
    grads = [None if p is None else torch.zeros_like(p) for p in table]
    ds, dr = torch.randn_like(sigma), torch.randn_like(rgb)
    trace = torch.zeros(2 * 160, dtype=torch.int64, device=DEV)
    for it in range(3):
        if it == 2:
            lib.snerfdbg_set_wgrad_trace(trace.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.mlp_backward(block.desc, table, packed, o, d, vd, z, sigma, rgb, ds, dr, grads, ws, flags)
        e1.record()
        torch.cuda.synchronize()
    lib.snerfdbg_set_wgrad_trace(None)
    t = trace.cpu().numpy().reshape(-1, 2)
    print(f'== {name}: backward {e0.elapsed_time(e1):.3f} ms')
    jobs = {}
    for cta in range(148):
        jobs.setdefault(int(t[cta, 0]), []).append(int(t[cta, 1]))
    for j, v in sorted(jobs.items()):
        print(f'  job {j:2d}: {len(v):3d} CTAs, cycles min {min(v)} max {max(v)}  -> work {sum(v) / 1e6:.2f} Mcyc')
        if max(v) > 1.1 * min(v):      # uneven inside one job: every CTA (in CTA order), kilo-cycles
            print('           per CTA:', [c // 1000 for c in v])
