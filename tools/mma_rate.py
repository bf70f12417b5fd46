"""Issue / execution rate of back-to-back tcgen05.mma (cta_group::1, SS operands) measured with clock64."""
import os as _os; _os.environ['SNERF_B200_DEBUG_LIB'] = '1'   # snerfdbg_* entry points live in libsimplenerf_b200_dbg.so (build.py --debug)
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from simplenerf_b200 import _lib
lib = _lib.load()
lib.snerfdbg_probe.restype = C.c_int
lib.snerfdbg_probe.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32,
                               C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
def idesc(m, n, a_mn=0, b_mn=0):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)
dev = 'cuda:0'
a = torch.zeros(65536, dtype=torch.uint8, device=dev); b = torch.zeros(131072, dtype=torch.uint8, device=dev)
d = torch.zeros((128, 512), device=dev); tm = torch.zeros(2, dtype=torch.int64, device=dev)
lib.snerfdbg_set_probe_pattern.argtypes = [C.c_int, C.c_int]
for chunk, waits in ((0, 0), (4, 0), (4, 1), (4, 2), (8, 2), (1, 0)):
  lib.snerfdbg_set_probe_pattern(chunk, waits)
  print(f'-- commit every {chunk} MMAs, {waits} ready-barrier waits per chunk')
  for n in (256,):
    for count in (64, 256):
        for same in (False,):
              ops = []
              for i in range(count):
                  k = i % 4; pan = 0 if same else (i // 4) % 4
                  ops.append((16384 * pan + 32 * k, (32768 * pan if n == 256 else 16384 * pan) + 32 * k, 0, int(i > 0)))
              o = torch.tensor(np.array(ops, np.int64), dtype=torch.int64).to(torch.int32).to(dev)
              for rep in range(2):
                  rc = lib.snerfdbg_probe(a.data_ptr(), 65536, b.data_ptr(), 131072, d.data_ptr(), o.data_ptr(), count, 16, 1024, 16, 1024,
                                          idesc(128, n), (1 << 46) | (2 << 61), 512, None, tm.data_ptr())
                  assert rc == 0
                  torch.cuda.synchronize()
              t = tm.cpu().numpy()
              print(f'N={n:3d} count={count:3d} same_operand={same}: issue {t[0]/count:6.1f} cyc/MMA, complete {t[1]/count:6.1f} cyc/MMA', flush=True)
