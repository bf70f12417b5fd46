"""Per-SM shared -> global store throughput: bulk copies (the stash writers) against per-thread vector stores (store_probe.cu)."""
import os as _os; _os.environ['SNERF_B200_DEBUG_LIB'] = '1'
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from simplenerf_b200 import _lib
lib = _lib.load()
lib.snerfdbg_store_probe.restype = C.c_int
lib.snerfdbg_store_probe.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
dev = 'cuda:0'
reps = 400
names = {0: 'bulk 64 KB', 1: 'st.global.v4 (512 threads)', 2: 'half bulk + half st.global', 3: 'bulk 4 x 16 KB'}
print('| mode | in flight | window per SM | SMs | B/clk/SM | GB/s total |')
print('|---|---|---|---|---|---|')
for grid in (148, 74):
    for window, wname in ((128 << 10, '128 KB (L2)'), (64 << 20, '64 MB (HBM)')):
        buf = torch.empty(grid * window, dtype=torch.uint8, device=dev)
        cyc = torch.zeros(grid, dtype=torch.int64, device=dev)
        for mode in (0, 3, 1, 2):
            for depth in ((1, 2, 4) if mode in (0, 3) else (1,)):
                for it in range(2):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    rc = lib.snerfdbg_store_probe(buf.data_ptr(), window, reps, mode, depth, grid, cyc.data_ptr(), None)
                    e1.record()
                    torch.cuda.synchronize()
                    assert rc == 0, lib.snerf_last_error()
                ms = e0.elapsed_time(e1)
                c = float(cyc.float().mean())
                print(f'| {names[mode]} | {depth} | {wname} | {grid} | {reps * 65536 / c:.1f} | {grid * reps * 65536 / ms / 1e6:.0f} |')
        del buf
