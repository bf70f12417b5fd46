"""Compile the CUDA sources into simplenerf_b200/libsimplenerf_b200.so (in-tree, sm_100a only)."""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'libsimplenerf_b200.so')
LIB_DBG = os.path.join(PKG, 'libsimplenerf_b200_dbg.so')
SOURCES = ['api.cu', 'sampling.cu', 'composite.cu', 'mlp_simt.cu', 'mlp_tc.cu', 'mlp_tc_bwd.cu', 'vis_tc.cu', 'adam.cu', 'raygen.cu', 'losses.cu', 'gather.cu']
# developer library: the product sources compiled with -DSNERF_DEBUG (snerfdbg_* entry points: clock64 traces, stage
# switches, descriptor probe) plus the stand-alone probe kernels -- kept out of the product library
DEBUG_SOURCES = SOURCES + ['tmem_bench.cu', 'pair_probe.cu', 'store_probe.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def _stale(lib: str) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG, '..', 'include', 'simplenerf_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    """nvcc cross-compiles without a GPU; returns the path of the shared library (debug=True: the developer library)."""
    lib = LIB_DBG if debug else LIB
    if not force and not _stale(lib):
        return lib
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-DSNERF_DEBUG'] if debug else []) + (['-Xptxas', '-v'] if verbose else []) + \
        (DEBUG_SOURCES if debug else SOURCES) + ['-o', lib]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'nvcc failed:\n{res.stdout}\n{res.stderr}')
    if verbose:
        print(res.stderr)
    return lib


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, debug='--debug' in sys.argv))
