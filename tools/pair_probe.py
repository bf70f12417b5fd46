"""Probe of the cta_group::2 (CTA pair) MMA: D[256 x N] = A[256 x 64] B[N x 64]^T with exact small integers.
Checks the operand / accumulator split between the two CTAs that the chain kernels assume, and times 256 pair MMAs."""
import os as _os; _os.environ['SNERF_B200_DEBUG_LIB'] = '1'   # snerfdbg_* entry points live in libsimplenerf_b200_dbg.so (build.py --debug)
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from simplenerf_b200 import _lib
from tools.tc_probe_util import panel_image  # noqa: E402
lib = _lib.load()
lib.snerfdbg_pair_probe.restype = C.c_int
lib.snerfdbg_pair_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
dev = 'cuda:0'
rng = np.random.default_rng(0)
for mode in (0,):   # mode 1 (peer copy completing on the leader barrier) never completes on B200: the relay is the protocol
    for n in (256, 128):
        A = rng.integers(-3, 4, (256, 64)).astype(np.float32)
        B = rng.integers(-3, 4, (n, 64)).astype(np.float32)
        a = torch.frombuffer(bytearray(panel_image(A)), dtype=torch.uint8).to(dev)
        b = torch.frombuffer(bytearray(panel_image(B)), dtype=torch.uint8).to(dev)
        d = torch.zeros((256, n), device=dev)
        rc = lib.snerfdbg_pair_probe(a.data_ptr(), b.data_ptr(), d.data_ptr(), n, mode, None, None)
        assert rc == 0, lib.snerf_last_error()
        torch.cuda.synchronize()
        err = np.abs(d.cpu().numpy() - A @ B.T).max()
        print(f'mode {mode} (peer copy completes on {"the leader barrier" if mode else "its own barrier + relay"}) N={n}: max abs err {err:g}', flush=True)
tm = torch.zeros(2, dtype=torch.int64, device=dev)
for n in (256, 128):
    A = np.zeros((256, 64), np.float32); B = np.zeros((n, 64), np.float32)
    a = torch.frombuffer(bytearray(panel_image(A)), dtype=torch.uint8).to(dev)
    b = torch.frombuffer(bytearray(panel_image(B)), dtype=torch.uint8).to(dev)
    d = torch.zeros((256, n), device=dev)
    for mode, what in ((0, 'back to back'), (2, 'commit per 4'), (6, 'commit + cluster-scope wait per 4'), (10, 'commit + cta-scope wait per 4'),
                       (18, 'commit + cluster-scope test_wait per 4')):
        for _ in range(2):
            lib.snerfdbg_pair_probe(a.data_ptr(), b.data_ptr(), d.data_ptr(), n, mode, tm.data_ptr(), None)
            torch.cuda.synchronize()
        t = tm.cpu().numpy()
        print(f'pair MMA M=256 N={n} K=16, {what}: issue {t[0] / 256:.1f} cyc/MMA, complete {t[1] / 256:.1f} cyc/MMA', flush=True)
