"""Round 2: where does the fused backward launch spend its time?  Times the whole launch, the dgrad chain alone and the
weight-gradient jobs alone (debug bits 8 / 16), the latter also without MMAs (2), without reducer work (4) and without both (6),
and prints the per-job cycle counts of the weight-gradient CTAs.  Run under SNERF_BWD_RING / SNERF_BWD_DGRAD_PAIRS settings."""
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
n_rays = 4096
print('ring', os.environ.get('SNERF_BWD_RING'), 'pairs', os.environ.get('SNERF_BWD_DGRAD_PAIRS'), 'spread', os.environ.get('SNERF_BWD_SPREAD'))
for name, key, s in (('fine', 'fine_mlp', 192), ('pts-aug', None, 64)):
    model_cfg = synthetic.make_configs('simplenerf')['model']
    cfg = model_cfg[key] if key else model_cfg['points_augmentation']['coarse_mlp']
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
    for label, bits in (('whole launch', 0), ('dgrad chain alone', 8), ('wgrad jobs alone', 16), ('wgrad alone, no MMAs', 16 | 2),
                        ('wgrad alone, no reducers', 16 | 4), ('wgrad alone, loads only', 16 | 6),
                        ('wgrad alone, loads only, no B (act) copies', 16 | 6 | 32), ('wgrad alone, loads only, no A (dY) copies', 16 | 6 | 64),
                        ('wgrad alone, no copies at all', 16 | 6 | 32 | 64)):
        lib.snerfdbg_set_wgrad_debug(bits)
        trace.zero_()
        best = 1e9
        for it in range(4):
            if it == 3:
                lib.snerfdbg_set_wgrad_trace(trace.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.mlp_backward(block.desc, table, packed, o, d, vd, z, sigma, rgb, ds, dr, grads, ws, flags)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        lib.snerfdbg_set_wgrad_trace(None)
        print(f'== {name}: {label}: {best:.3f} ms')
        if bits != 8:
            t = trace.cpu().numpy().reshape(-1, 2)
            jobs = {}
            for cta in range(148):
                if t[cta, 1] > 0:
                    jobs.setdefault(int(t[cta, 0]), []).append(int(t[cta, 1]))
            print('   ' + ' | '.join(f'j{j}:{len(v)}x{max(v) / 1e6:.2f}M' for j, v in sorted(jobs.items())))
    lib.snerfdbg_set_wgrad_debug(0)
