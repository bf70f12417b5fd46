"""ctypes binding of include/simplenerf_b200.h.  There is no fallback: if the shared library is
missing or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, 'libsimplenerf_b200.so')
# the same sources + the probe kernels and the snerfdbg_* entry points (compiled with -DSNERF_DEBUG by `build.py --debug`); the
# developer tools under tools/ select it with SNERF_B200_DEBUG_LIB=1, the product never loads it
LIB_DBG_PATH = os.path.join(PKG, 'libsimplenerf_b200_dbg.so')

P_COUNT = 24
P_HEAD_W, P_HEAD_B, P_FEAT_W, P_FEAT_B, P_VIEW_W, P_VIEW_B, P_RGB_W, P_RGB_B = 16, 17, 18, 19, 20, 21, 22, 23
FLAG_NDC, FLAG_WHITE_BKGD, FLAG_LINDISP, FLAG_SAVE_FOR_BWD, FLAG_PRECISE, FLAG_VIS_GRAD, FLAG_VIS_HEAD = 1, 2, 4, 8, 16, 32, 64


class MlpDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('depth', 'width', 'skip_layer', 'pts_degree', 'trunk_degree', 'view_degree',
                                         'view_width', 'head_out')]


class LossStream(C.Structure):
    _fields_ = [('pred', C.c_void_p), ('target', C.c_void_p), ('mask', C.c_void_p), ('grad', C.c_void_p),
                ('channels', C.c_int32), ('weight', C.c_float), ('kind', C.c_int32)]


LOSS_SQUARED, LOSS_ABSOLUTE, LOSS_PRIOR_SHORTFALL = 0, 1, 2


class ReprojArgs(C.Structure):
    _fields_ = [('depth_main', C.c_void_p), ('depth_other', C.c_void_p * 4), ('grad_main', C.c_void_p), ('grad_other', C.c_void_p * 4),
                ('weight', C.c_float * 4), ('n_others', C.c_int32), ('rays_o', C.c_void_p), ('rays_d', C.c_void_p),
                ('pixel_id', C.c_void_p), ('mask_nerf', C.c_void_p), ('images', C.c_void_p), ('proj', C.c_void_p),
                ('origins', C.c_void_p), ('closest', C.c_void_p), ('n_views', C.c_int32), ('height', C.c_int32),
                ('width', C.c_int32), ('half_patch', C.c_int32), ('rmse_threshold', C.c_float), ('flags', C.c_uint32)]


class GatherTable(C.Structure):
    _fields_ = [('src', C.c_void_p), ('dst', C.c_void_p), ('mask', C.c_void_p), ('row_bytes', C.c_int32), ('fill_bits', C.c_uint32)]


GATHER_MAX_TABLES = 24
LOSS_MAX_STREAMS = 8
REPROJ_MAX_OTHERS = 4
REPROJ_SYMMETRIC = 1
_fp = C.c_void_p   # device pointers travel as integers
_SIGNATURES = {
    'snerf_abi_version': (C.c_int, []),
    'snerf_last_error': (C.c_char_p, []),
    'snerf_has_tensor_path': (C.c_int, []),
    'snerf_sample_coarse': (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_uint32, _fp]),
    'snerf_sample_fine': (C.c_int, [_fp, _fp, _fp, C.c_int, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    'snerf_composite_forward': (C.c_int, [_fp] * 15 + [C.c_int, C.c_int, C.c_uint32, _fp]),
    'snerf_composite_backward': (C.c_int, [_fp] * 17 + [C.c_int, C.c_int, C.c_uint32, _fp]),
    'snerf_mlp_workspace_bytes': (C.c_size_t, [C.POINTER(MlpDesc), C.c_int, C.c_int, C.c_uint32]),
    'snerf_render_workspace_bytes': (C.c_size_t, [C.POINTER(MlpDesc), C.c_int, C.c_int, C.c_uint32]),
    'snerf_render_forward': (C.c_int, [C.POINTER(MlpDesc), C.POINTER(_fp), _fp] + [_fp] * 14 + [_fp, C.c_size_t, C.c_int, C.c_int, C.c_uint32, _fp]),
    'snerf_packed_weights_bytes': (C.c_size_t, [C.POINTER(MlpDesc)]),
    'snerf_pack_weights': (C.c_int, [C.POINTER(MlpDesc), C.POINTER(_fp), _fp, _fp]),
    'snerf_mlp_forward': (C.c_int, [C.POINTER(MlpDesc), C.POINTER(_fp), _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                                    C.c_size_t, C.c_int, C.c_int, C.c_uint32, _fp]),
    'snerf_mlp_backward': (C.c_int, [C.POINTER(MlpDesc), C.POINTER(_fp), _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                                     C.POINTER(_fp), _fp, C.c_size_t, C.c_int, C.c_int, C.c_uint32, _fp]),
    'snerf_generate_rays': (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, _fp, _fp, _fp]),
    'snerf_postprocess_frame': (C.c_int, [_fp, _fp, C.POINTER(_fp), C.c_int, C.c_longlong, _fp]),
    'snerf_adam_step': (C.c_int, [C.POINTER(_fp), C.POINTER(_fp), C.POINTER(_fp), C.POINTER(_fp), C.POINTER(C.c_longlong), C.c_int,
                                  C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, _fp]),
    'snerf_ray_losses_workspace_bytes': (C.c_size_t, []),
    'snerf_ray_losses_forward': (C.c_int, [C.POINTER(LossStream), C.c_int, C.c_int, _fp, _fp, _fp, C.c_size_t, _fp]),
    'snerf_ray_losses_backward': (C.c_int, [C.POINTER(LossStream), C.c_int, C.c_int, _fp, _fp, _fp]),
    'snerf_ray_loss_maps': (C.c_int, [C.POINTER(LossStream), C.c_int, C.c_int, _fp]),
    'snerf_reprojection_losses_forward': (C.c_int, [C.POINTER(ReprojArgs), C.c_int, _fp, _fp, _fp, _fp, C.c_size_t, _fp]),
    'snerf_reprojection_losses_backward': (C.c_int, [C.POINTER(ReprojArgs), C.c_int, _fp, _fp, _fp, _fp]),
    'snerf_gather_rows': (C.c_int, [C.POINTER(GatherTable), C.c_int, _fp, C.c_int, _fp]),
    'snerf_visibility_workspace_bytes': (C.c_size_t, [C.POINTER(MlpDesc), C.c_int, C.c_int, C.c_int]),
    'snerf_visibility_forward': (C.c_int, [C.POINTER(MlpDesc), C.POINTER(_fp)] + [_fp] * 8 + [C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_uint32, _fp]),
    'snerf_visibility_backward': (C.c_int, [C.POINTER(MlpDesc), C.POINTER(_fp)] + [_fp] * 9 + [C.POINTER(_fp), _fp, C.c_size_t,
                                            C.c_int, C.c_int, C.c_int, C.c_uint32, _fp]),
    'snerf_visibility2_composite_forward': (C.c_int, [_fp] * 4 + [C.c_int, C.c_int, C.c_int, _fp]),
    'snerf_visibility2_composite_backward': (C.c_int, [_fp] * 8 + [C.c_int, C.c_int, C.c_int, _fp]),
    'snerf_tensor_selftest': (C.c_int, [C.POINTER(C.c_float), _fp]),
    'snerf_fill_random': (C.c_int, [_fp, C.c_longlong, C.c_int, C.c_float, C.c_uint64, C.c_uint64, _fp]),
    'snerf_sample_coarse_rng': (C.c_int, [_fp, _fp, _fp, C.c_uint64, C.c_uint64, _fp, C.c_int, C.c_int, C.c_uint32, _fp]),
    'snerf_sample_fine_rng': (C.c_int, [_fp, _fp, C.c_uint64, C.c_uint64, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    'snerf_mlp_forward_rng': (C.c_int, [C.POINTER(MlpDesc), C.POINTER(_fp), _fp, _fp, _fp, _fp, _fp, C.c_float, C.c_uint64, C.c_uint64,
                                        _fp, _fp, _fp, C.c_size_t, C.c_int, C.c_int, C.c_uint32, _fp]),
    'snerf_set_backward_split_event': (None, [_fp]),
}
EXPORTS = tuple(_SIGNATURES)
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        path = LIB_DBG_PATH if os.environ.get('SNERF_B200_DEBUG_LIB') == '1' else LIB_PATH
        path = os.environ.get('SNERF_B200_LIB_AB', path)     # developer A/B runs (tools/ab.sh): another build of the same sources
        if not os.path.exists(path):
            raise RuntimeError(f'{path} is missing: build it with `python -m simplenerf_b200.build` '
                               '(there is no CPU or PyTorch fallback for this path)')
        lib = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.snerf_abi_version() != 1:
            raise RuntimeError('libsimplenerf_b200.so: ABI version mismatch')
        _lib = lib
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        raise RuntimeError(f'{what} failed (status {status}): {load().snerf_last_error().decode()}')
