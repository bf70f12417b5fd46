"""Swizzled bf16 panel images for the descriptor probes."""
import os as _os; _os.environ['SNERF_B200_DEBUG_LIB'] = '1'   # snerfdbg_* entry points live in libsimplenerf_b200_dbg.so (build.py --debug)
import numpy as np


def bf16_bits(x):
    return (x.astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)


def panel_image(mat):
    """mat [rows, 64] (float) -> 128-byte-swizzled bf16 image bytes (rows*128)."""
    rows = mat.shape[0]
    bits = bf16_bits(mat).reshape(rows, 8, 8)
    out = np.zeros((rows, 8, 8), np.uint16)
    for r in range(rows):
        for c in range(8):
            out[r, c ^ (r & 7)] = bits[r, c]
    return out.tobytes()
