"""Probe of the tcgen05 shared-memory / instruction descriptors used by simplenerf_b200/csrc/mlp_tc.cu.
Runs small exact-integer GEMMs through the debug entry `snerfdbg_probe` and compares with numpy.
Usage (on a B200):  python tools/tc_probe.py"""
import os as _os; _os.environ['SNERF_B200_DEBUG_LIB'] = '1'   # snerfdbg_* entry points live in libsimplenerf_b200_dbg.so (build.py --debug)
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simplenerf_b200 import _lib  # noqa: E402

lib = _lib.load()
lib.snerfdbg_probe.restype = C.c_int
lib.snerfdbg_probe.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_uint32,
                               C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
VERSION, SW128 = 1 << 46, 2 << 61


def bf16_bits(x):
    return (x.astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)


def panel_image(mat):
    """mat [rows, 64] (float) -> swizzled bf16 image bytes (rows*128)."""
    rows = mat.shape[0]
    bits = bf16_bits(mat).reshape(rows, 8, 8)
    out = np.zeros((rows, 8, 8), np.uint16)
    for r in range(rows):
        for c in range(8):
            out[r, c ^ (r & 7)] = bits[r, c]
    return out.tobytes()


def panels(mat):
    """[rows, 64*k] -> k consecutive panel images."""
    return b''.join(panel_image(mat[:, 64 * j:64 * (j + 1)]) for j in range(mat.shape[1] // 64))


def idesc(m, n, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def run(a_img, b_img, ops, a_lbo, a_sbo, b_lbo, b_sbo, idsc, n_cols, bits=VERSION | SW128):
    dev = 'cuda:0'
    a = torch.frombuffer(bytearray(a_img), dtype=torch.uint8).to(dev)
    b = torch.frombuffer(bytearray(b_img), dtype=torch.uint8).to(dev)
    o = torch.tensor(np.array(ops, np.uint32).reshape(-1, 4).astype(np.int64), dtype=torch.int64).to(torch.int32).to(dev)
    d = torch.zeros((128, n_cols), device=dev)
    rc = lib.snerfdbg_probe(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), d.data_ptr(), o.data_ptr(), len(ops), a_lbo, a_sbo,
                            b_lbo, b_sbo, idsc, bits, n_cols, None, None)
    assert rc == 0, lib.snerf_last_error()
    torch.cuda.synchronize()
    return d.cpu().numpy()


def report(name, got, want):
    err = np.abs(got - want).max()
    print(f'{name}: max abs err {err:g}  (|want| max {np.abs(want).max():g}, exact={err == 0})', flush=True)
    return err == 0


rng = np.random.default_rng(0)
ok = True
# T0: K-major, one 64-wide K panel
A = rng.integers(-3, 4, (128, 64)).astype(np.float32)
B = rng.integers(-3, 4, (256, 64)).astype(np.float32)
ops = [(32 * k, 32 * k, 0, int(k > 0)) for k in range(4)]
ok &= report('T0 K-major M128 N256 K64', run(panels(A), panels(B), ops, 16, 1024, 16, 1024, idesc(128, 256, 0, 0), 256), A @ B.T)
# T1: K-major, K = 128 (two panels / two weight chunks), second accumulator half
A = rng.integers(-3, 4, (128, 128)).astype(np.float32)
B = rng.integers(-3, 4, (256, 128)).astype(np.float32)
ops = [(16384 * j + 32 * k, 32768 * j + 32 * k, 256, int(j + k > 0)) for j in range(2) for k in range(4)]
got = run(panels(A), panels(B), ops, 16, 1024, 16, 1024, idesc(128, 256, 0, 0), 512)
ok &= report('T1 K-major K128, D at col 256', got[:, 256:], A @ B.T)
# T2: N = 128
B = rng.integers(-3, 4, (128, 64)).astype(np.float32)
A = rng.integers(-3, 4, (128, 64)).astype(np.float32)
ops = [(32 * k, 32 * k, 0, int(k > 0)) for k in range(4)]
ok &= report('T2 K-major N128', run(panels(A), panels(B), ops, 16, 1024, 16, 1024, idesc(128, 128, 0, 0), 128), A @ B.T)
# T3: MN-major both (wgrad): D[m][n] = sum_p Y[p][m] X[p][n];  Y [128 pts,128], X [128 pts,256]
Y = rng.integers(-3, 4, (128, 128)).astype(np.float32)
X = rng.integers(-3, 4, (128, 256)).astype(np.float32)
want = Y.T @ X
ops = [(2048 * k, 2048 * k, 0, int(k > 0)) for k in range(8)]
for (lbo, sbo) in ((16384, 1024), (1024, 16384)):
    got = run(panels(Y), panels(X), ops, lbo, sbo, lbo, sbo, idesc(128, 256, 1, 1), 256)
    good = report(f'T3 MN-major lbo={lbo} sbo={sbo}', got, want)
    if (lbo, sbo) == (16384, 1024):
        ok &= good
# T4: MN-major with 64-point half panels (panel stride 8 KiB), K = 64 points
img_y = b''.join(panel_image(Y[:64, 64 * j:64 * (j + 1)]) for j in range(2))
img_x = b''.join(panel_image(X[:64, 64 * j:64 * (j + 1)]) for j in range(4))
ops = [(2048 * k, 2048 * k, 0, int(k > 0)) for k in range(4)]
ok &= report('T4 MN-major half panels', run(img_y, img_x, ops, 8192, 1024, 8192, 1024, idesc(128, 256, 1, 1), 256),
             Y[:64].T @ X[:64])
# T5: MN-major, no swizzle ("interleave"): 64-point blocks stored chunk-major [8 chunks][64 points][16 B], a layout an
# epilogue could write straight from registers with coalesced 16-byte stores.  Core matrix = 8 points x 16 B contiguous.
# (Tried for the stash: the descriptor works, but per-thread global stores from the epilogue warps measured 6-17 % slower
# than the bulk shared->global copies of swizzled panels, so the kernels keep the latter.)
def block_image(mat):
    """mat [64 points, 64 cols] -> [chunk][point][8 bf16]"""
    bits = bf16_bits(mat).reshape(64, 8, 8)
    return np.ascontiguousarray(bits.transpose(1, 0, 2)).tobytes()


img_y = b''.join(block_image(Y[:64, 64 * j:64 * (j + 1)]) for j in range(2))
img_x = b''.join(block_image(X[:64, 64 * j:64 * (j + 1)]) for j in range(4))
ops = [(256 * k, 256 * k, 0, int(k > 0)) for k in range(4)]
for (lbo, sbo) in ((128, 1024), (1024, 128)):
    got = run(img_y, img_x, ops, lbo, sbo, lbo, sbo, idesc(128, 256, 1, 1), 256, bits=VERSION)
    report(f'T5 MN-major interleave lbo={lbo} sbo={sbo}', got, Y[:64].T @ X[:64])
print('ALL OK' if ok else 'SOME FAILED')
