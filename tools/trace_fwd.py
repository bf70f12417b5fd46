"""Timeline of one tile of the forward chain kernel (CTA 0, third tile): clock64 stamps of the MMA issuer and of one
epilogue warp.  Debug tool for the pipeline analysis in profiles/."""
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
save = '--save' in sys.argv
cfg = synthetic.make_configs('simplenerf')['model']['coarse_mlp']
block = MlpBlock(cfg).to(DEV)
table = [None if p is None else p.detach() for p in block.param_table()]
packed = block.packed(table)
n_rays, s = 4096, 64
b = synthetic.make_ray_batch('llff', n_rays, 3)
o, d, vd = b['rays_o_ndc'].to(DEV), b['rays_d_ndc'].to(DEV), b['view_dirs'].to(DEV)
z = torch.sort(torch.rand(n_rays, s, device=DEV), -1)[0].contiguous()
flags = FLAG_SAVE_FOR_BWD if save else 0
ws = torch.empty(ops.mlp_workspace_bytes(block.desc, n_rays, s, flags), dtype=torch.uint8, device=DEV)
trace = torch.zeros(1024, dtype=torch.int64, device=DEV)
for _ in range(2):
    ops.mlp_forward(block.desc, table, packed, o, d, vd, z, None, ws, flags)
lib.snerfdbg_set_trace.argtypes = [ctypes.c_void_p]
lib.snerfdbg_set_fwd_debug.argtypes = [ctypes.c_int]
dbg = [int(a[6:]) for a in sys.argv if a.startswith('--dbg=')]
lib.snerfdbg_set_fwd_debug(dbg[0] if dbg else 0)
lib.snerfdbg_set_trace(trace.data_ptr())
ops.mlp_forward(block.desc, table, packed, o, d, vd, z, None, ws, flags)
torch.cuda.synchronize()
lib.snerfdbg_set_trace(None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.mlp_forward(block.desc, table, packed, o, d, vd, z, None, ws, flags)
e1.record(); torch.cuda.synchronize()
print(f'forward kernel (4096 rays x 64, save={save}): {e0.elapsed_time(e1) / 5 * 1e3:.1f} us')
t = trace.cpu().numpy()
t0 = t[0]   # issuer: slot 0, step 0 of the pair's third tile: operands ready
rel = lambda v: int(v - t0) if v else None
print('slot step | issuer: operands ready, all MMAs issued | epilogue warp 6: accumulator complete, handed over')
for s_ in range(10):
    for x in range(2):
        e = t[(x * 16 + s_) * 16: (x * 16 + s_) * 16 + 16]
        print(f'{x} {s_:2d} | {rel(e[0])} {rel(e[1])} | {rel(e[2])} {rel(e[3])} | weights seen {[rel(v) for v in e[4:9]]} chunk issued {[rel(v) for v in e[9:14]]}')
print('epilogue warp 6 detail: slot step | after acc wait | unit 0..3 stored | fenced | arrived   (relative to the acc wait)')
for s_ in range(10):
    for x in range(2):
        e = t[(x * 16 + s_) * 16: (x * 16 + s_) * 16 + 16]
        d = t[512 + (x * 16 + s_) * 8: 512 + (x * 16 + s_) * 8 + 8]
        r2 = lambda v: int(v - e[2]) if v else None
        print(f'{x} {s_:2d} | 0 | {[r2(v) for v in d[1:5]]} | {r2(d[5])} | {r2(e[3])} | unit 0: loaded {r2(d[6])} computed {r2(d[7])}')
