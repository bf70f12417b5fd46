"""Timeline of one tile of the forward chain kernel (CTA 0, third tile): clock64 stamps of the MMA issuer and of one
epilogue warp.  Debug tool for the pipeline analysis in profiles/."""
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
print(f'forward kernel (4096 rays x 64): {e0.elapsed_time(e1) / 5 * 1e3:.1f} us')
t = trace.cpu().numpy()
t0 = t[256]   # epilogue: accumulator of step 0 complete
rel = lambda v: int(v - t0) if v else None
print('EPI  step | acc done | tmem-ld done p0..p3 | handed p0..p3 | panel1: math done, stores issued, fence done')
for s_ in range(10):
    e = t[256 + s_ * 16: 256 + s_ * 16 + 16]
    print(f'{s_:2d} | {rel(e[0])} | {[rel(x) for x in e[1:5]]} | {[rel(x) for x in e[5:9]]} | {[rel(x) for x in e[9:12]]}')
print('MMA  step | A panel available c0..c4 | weight chunk landed c0..c4 | chunk issued c0..c4')
for s_ in range(10):
    m = t[s_ * 16: s_ * 16 + 16]
    print(f'{s_:2d} | {[rel(x) for x in m[0:5]]} | {[rel(x) for x in m[5:10]]} | {[rel(x) for x in m[10:15]]}')
print('all epilogue warps (warp, q, hf): handed p0..p3 of steps 1 and 2')
for w in range(8):
    v = t[512 + w * 8: 512 + w * 8 + 8]
    print(f'warp {w + 2} q={(w + 2) & 3} hf={w >> 2}: {[rel(x) for x in v]}')
