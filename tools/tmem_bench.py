"""TMEM -> register read throughput (tcgen05.ld) per SM: shapes, warp counts, with / without concurrent MMAs."""
import os as _os; _os.environ['SNERF_B200_DEBUG_LIB'] = '1'   # snerfdbg_* entry points live in libsimplenerf_b200_dbg.so (build.py --debug)
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from simplenerf_b200 import _lib
lib = _lib.load()
lib.snerfdbg_tmem_bench.restype = C.c_int
lib.snerfdbg_tmem_bench.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
out = torch.zeros(64, dtype=torch.int64, device='cuda:0')
reps = 256
print('| shape | warps | pipelined | MMA running | cycles / load (slowest warp) | bytes/clk/SM | cycles / MMA (M128 N256 K16) while the loads run |')
print('|---|---|---|---|---|---|---|')
for shape, bytes_per in ((32, 4096), (16, 2048), (256, 4096)):
    for warps in (1, 4, 8, 16):
        for pipelined in ((0, 1) if shape == 32 else (0,)):
            for mma in (0, 1):
                out.zero_()
                for _ in range(2):
                    rc = lib.snerfdbg_tmem_bench(out.data_ptr(), warps, reps, shape, mma, pipelined, None)
                    assert rc == 0
                    torch.cuda.synchronize()
                t = out.cpu().numpy()
                worst = max(int(t[w]) for w in range(1, 1 + warps))
                print(f'| {"32x32b.x%d" % shape if shape != 256 else "16x256b.x8"} | {warps} | {pipelined} | {mma} | {worst / reps:.1f} | '
                      f'{warps * reps * bytes_per / worst:.1f} | {(t[61] / max(t[63], 1)) if mma else float("nan"):.1f} |', flush=True)
