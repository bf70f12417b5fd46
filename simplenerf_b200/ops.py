"""Tensor-level wrappers over the C ABI (include/simplenerf_b200.h).

torch is plumbing here: it owns device memory and the current CUDA stream; all arithmetic happens
in the kernels of libsimplenerf_b200.so.  Every wrapper raises if a tensor is not a contiguous
fp32 CUDA tensor -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import FLAG_LINDISP, FLAG_NDC, FLAG_PRECISE, FLAG_SAVE_FOR_BWD, FLAG_VIS_GRAD, FLAG_WHITE_BKGD, MlpDesc, P_COUNT  # noqa: F401


# torch.cuda.current_device() without its lazy-init checks (the tensors checked below are CUDA tensors, so CUDA is up): _ptr runs
# ~650 times per training step
_current_device = getattr(torch._C, '_cuda_getDevice', torch.cuda.current_device)


def _ptr(t: Optional[torch.Tensor], dtype=torch.float32) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('simplenerf_b200 kernels need CUDA tensors (no CPU fallback exists)')
    if t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError(f'expected a contiguous {dtype} tensor, got {t.dtype}, contiguous={t.is_contiguous()}')
    if t.device.index != _current_device():
        # the launch goes to the CURRENT device's stream (_stream): a tensor of another device would hand the kernel a
        # foreign pointer.  The drop-in's forward / backward select the batch's device; direct callers must do the same.
        raise RuntimeError(f'tensor lives on cuda:{t.device.index} but the current device is cuda:{torch.cuda.current_device()}: '
                           'wrap the call in `with torch.cuda.device(tensor.device):`')
    return t.data_ptr()


# kernels launched by this process through the C ABI (bench.py reports it as gpu_launches)
LAUNCHES = {'count': 0}
# optional hook: a callable (name) -> context manager, used by bench.py to time the MLP kernels with CUDA events
TIMER = {'hook': None}


class _NoTimer:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _timed(name: str):
    hook = TIMER['hook']
    return hook(name) if hook is not None else _NoTimer()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


# --------------------------------------------------------------------------------------------
class RngDraw:
    """One draw of the in-kernel generator (Philox4x32-10 keyed by `seed`, counter = (element / 4, `offset`)): passed where a
    tensor of random numbers would go, the consuming kernel draws the numbers itself (snerf_*_rng).  `materialize` writes the
    very same numbers to memory (snerf_fill_random) for the code paths that need a tensor."""

    def __init__(self, seed: int, offset: int, scale: float = 1.0):
        self.seed, self.offset, self.scale = int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1), float(scale)

    def materialize(self, shape, device, normal: bool) -> torch.Tensor:
        out = torch.empty(shape, device=device, dtype=torch.float32)
        LAUNCHES['count'] += 1
        with torch.cuda.device(out.device):
            _lib.check(_lib.load().snerf_fill_random(_ptr(out), out.numel(), int(normal), self.scale, self.seed, self.offset, _stream()),
                       'snerf_fill_random')
        return out


def sample_coarse(near: torch.Tensor, far: torch.Tensor, t_vals: torch.Tensor, t_rand, lindisp: bool = False) -> torch.Tensor:
    """get_z_vals_coarse (reference :272-302).  near/far [N,1] or [N]; t_vals [S]; t_rand [N,S], an RngDraw (drawn in the
    kernel) or None."""
    n, s = near.shape[0], t_vals.shape[0]
    near, far, t_vals = _f32(near).reshape(-1), _f32(far).reshape(-1), _f32(t_vals)
    if isinstance(t_rand, RngDraw):
        z = torch.empty((n, s), device=near.device, dtype=torch.float32)
        LAUNCHES['count'] += 1
        _lib.check(_lib.load().snerf_sample_coarse_rng(_ptr(near), _ptr(far), _ptr(t_vals), t_rand.seed, t_rand.offset, _ptr(z), n, s,
                                                       FLAG_LINDISP if lindisp else 0, _stream()), 'snerf_sample_coarse_rng')
        return z
    if t_rand is not None:
        t_rand = _f32(t_rand)
        assert tuple(t_rand.shape) == (n, s)
    z = torch.empty((n, s), device=near.device, dtype=torch.float32)
    LAUNCHES['count'] += 1
    _lib.check(_lib.load().snerf_sample_coarse(_ptr(near), _ptr(far), _ptr(t_vals), _ptr(t_rand), _ptr(z), n, s,
                                               FLAG_LINDISP if lindisp else 0, _stream()), 'snerf_sample_coarse')
    return z


def sample_fine(z_coarse: torch.Tensor, weights_coarse: torch.Tensor, u: torch.Tensor, debug: bool = False):
    """get_z_vals_fine + sample_pdf (reference :304-361).  u: [N,n_new] random draws, or [n_new] = the
    deterministic linspace row shared by all rays, or (RngDraw, n_new): drawn in the kernel."""
    n, sc = z_coarse.shape
    if isinstance(u, tuple) and isinstance(u[0], RngDraw):
        draw, n_new = u
        if debug:
            return sample_fine(z_coarse, weights_coarse, draw.materialize((n, n_new), z_coarse.device, normal=False), debug=True)
        z_coarse, weights_coarse = _f32(z_coarse), _f32(weights_coarse)
        z_fine = torch.empty((n, sc + n_new), device=z_coarse.device, dtype=torch.float32)
        LAUNCHES['count'] += 1
        _lib.check(_lib.load().snerf_sample_fine_rng(_ptr(z_coarse), _ptr(weights_coarse), draw.seed, draw.offset, _ptr(z_fine), n, sc,
                                                     n_new, _stream()), 'snerf_sample_fine_rng')
        return z_fine
    z_coarse, weights_coarse, u = _f32(z_coarse), _f32(weights_coarse), _f32(u)
    n_new = u.shape[-1]
    u_stride = 0 if u.dim() == 1 else n_new
    if u.dim() == 2:
        assert u.shape[0] == n
    dev = z_coarse.device
    z_fine = torch.empty((n, sc + n_new), device=dev, dtype=torch.float32)
    dbg = {}
    if debug:
        dbg = dict(samples=torch.empty((n, n_new), device=dev), cdf=torch.empty((n, sc - 1), device=dev),
                   below=torch.empty((n, n_new), device=dev, dtype=torch.int32),
                   above=torch.empty((n, n_new), device=dev, dtype=torch.int32))
    LAUNCHES['count'] += 1
    _lib.check(_lib.load().snerf_sample_fine(
        _ptr(z_coarse), _ptr(weights_coarse), _ptr(u), u_stride, _ptr(z_fine), _ptr(dbg.get('samples')),
        _ptr(dbg.get('cdf')), _ptr(dbg.get('below'), torch.int32), _ptr(dbg.get('above'), torch.int32), n, sc, n_new,
        _stream()), 'snerf_sample_fine')
    return (z_fine, dbg) if debug else z_fine


# --------------------------------------------------------------------------------------------
PER_RAY = ('rgb', 'acc', 'depth', 'depth_var', 'depth_ndc', 'depth_var_ndc')
PER_SAMPLE = ('alpha', 'visibility', 'weights')


def composite_forward(sigma, rgb, z, rays_o, rays_d, rays_d_ndc, ndc: bool, white_bkgd: bool = False,
                      per_sample: Sequence[str] = PER_SAMPLE) -> Dict[str, torch.Tensor]:
    """volume_rendering (reference :430-483).  sigma, z [N,S]; rgb [N,S,3]."""
    n, s = z.shape
    dev = z.device
    out = {'rgb': torch.empty((n, 3), device=dev), 'acc': torch.empty(n, device=dev),
           'depth': torch.empty(n, device=dev), 'depth_var': torch.empty(n, device=dev)}
    if ndc:
        out['depth_ndc'] = torch.empty(n, device=dev)
        out['depth_var_ndc'] = torch.empty(n, device=dev)
    for k in per_sample:
        out[k] = torch.empty((n, s), device=dev)
    flags = (FLAG_NDC if ndc else 0) | (FLAG_WHITE_BKGD if white_bkgd else 0)
    LAUNCHES['count'] += 1
    _lib.check(_lib.load().snerf_composite_forward(
        _ptr(sigma), _ptr(rgb), _ptr(z), _ptr(rays_o), _ptr(rays_d), _ptr(rays_d_ndc) if ndc else None,
        _ptr(out['rgb']), _ptr(out['acc']), _ptr(out['depth']), _ptr(out['depth_var']), _ptr(out.get('depth_ndc')),
        _ptr(out.get('depth_var_ndc')), _ptr(out.get('alpha')), _ptr(out.get('visibility')), _ptr(out.get('weights')),
        n, s, flags, _stream()), 'snerf_composite_forward')
    return out


def composite_backward(sigma, rgb, z, rays_o, rays_d, rays_d_ndc, ndc: bool, white_bkgd: bool,
                       grads: Dict[str, Optional[torch.Tensor]]):
    """grads: optional incoming gradients keyed like composite_forward's outputs."""
    n, s = z.shape
    d_sigma = torch.empty((n, s), device=z.device)
    d_rgb = torch.empty((n, s, 3), device=z.device)
    g = {k: (None if v is None else _f32(v)) for k, v in grads.items()}
    flags = (FLAG_NDC if ndc else 0) | (FLAG_WHITE_BKGD if white_bkgd else 0)
    LAUNCHES['count'] += 1
    _lib.check(_lib.load().snerf_composite_backward(
        _ptr(sigma), _ptr(rgb), _ptr(z), _ptr(rays_o), _ptr(rays_d), _ptr(rays_d_ndc) if ndc else None,
        _ptr(g.get('rgb')), _ptr(g.get('acc')), _ptr(g.get('depth')), _ptr(g.get('depth_var')), _ptr(g.get('depth_ndc')),
        _ptr(g.get('depth_var_ndc')), _ptr(g.get('alpha')), _ptr(g.get('visibility')), _ptr(g.get('weights')),
        _ptr(d_sigma), _ptr(d_rgb), n, s, flags, _stream()), 'snerf_composite_backward')
    return d_sigma, d_rgb


# --------------------------------------------------------------------------------------------
def pointer_table(tensors: Sequence[Optional[torch.Tensor]]):
    arr = (C.c_void_p * P_COUNT)()
    for i, t in enumerate(tensors):
        arr[i] = _ptr(t)
    return arr


def mlp_workspace_bytes(desc: MlpDesc, n_rays: int, n_samples: int, flags: int) -> int:
    b = _lib.load().snerf_mlp_workspace_bytes(C.byref(desc), n_rays, n_samples, flags)
    if b == 0:
        raise RuntimeError(f'snerf_mlp_workspace_bytes: {_lib.load().snerf_last_error().decode()}')
    return b


def packed_weights_bytes(desc: MlpDesc) -> int:
    return _lib.load().snerf_packed_weights_bytes(C.byref(desc))


def pack_weights(desc: MlpDesc, params: Sequence[Optional[torch.Tensor]], packed: torch.Tensor) -> None:
    LAUNCHES['count'] += 1
    _lib.check(_lib.load().snerf_pack_weights(C.byref(desc), pointer_table(params), _ptr(packed, torch.uint8), _stream()),
               'snerf_pack_weights')


def mlp_forward(desc: MlpDesc, params, packed, rays_o, rays_d, view_dirs, z, noise, workspace, flags: int):
    n, s = z.shape
    sigma = torch.empty((n, s), device=z.device)
    rgb = torch.empty((n, s, 3), device=z.device)
    # tensor path: view-bias kernel (view-dependent MLPs) + the fused chain kernel; precise path: ~14 launches
    LAUNCHES['count'] += (14 if flags & FLAG_PRECISE else (2 if desc.view_width else 1))
    if isinstance(noise, RngDraw):
        if flags & FLAG_PRECISE:        # the fp32 path reads a tensor: the same numbers, written out first
            noise = noise.materialize((n * s,), z.device, normal=True)
        else:                           # tensor path: drawn in the sigma-head epilogue
            with _timed('mlp_forward'):
                _lib.check(_lib.load().snerf_mlp_forward_rng(
                    C.byref(desc), pointer_table(params), _ptr(packed, torch.uint8), _ptr(rays_o), _ptr(rays_d), _ptr(view_dirs),
                    _ptr(z), noise.scale, noise.seed, noise.offset, _ptr(sigma), _ptr(rgb), _ptr(workspace, torch.uint8),
                    workspace.numel(), n, s, flags, _stream()), 'snerf_mlp_forward_rng')
            return sigma, rgb
    with _timed('mlp_forward'):
        _lib.check(_lib.load().snerf_mlp_forward(
            C.byref(desc), pointer_table(params), _ptr(packed, torch.uint8), _ptr(rays_o), _ptr(rays_d), _ptr(view_dirs),
            _ptr(z), _ptr(noise), _ptr(sigma), _ptr(rgb), _ptr(workspace, torch.uint8), workspace.numel(), n, s, flags,
            _stream()), 'snerf_mlp_forward')
    return sigma, rgb


def render_forward(desc: MlpDesc, params, packed, pts_o, pts_d, view_dirs, z, rays_o, rays_d, ndc: bool, white_bkgd: bool,
                   want_weights: bool) -> Dict[str, torch.Tensor]:
    """Evaluation with the samples kept on chip: run_network + volume_rendering (reference :153-160, :430-483)
    in one pass -- sigma / rgb never reach HBM.  Returns the per-ray maps, `alpha` [N,S] and, on request, `weights` [N,S]."""
    n, s = z.shape
    dev = z.device
    out = {'rgb': torch.empty((n, 3), device=dev), 'acc': torch.empty(n, device=dev),
           'depth': torch.empty(n, device=dev), 'depth_var': torch.empty(n, device=dev)}
    if ndc:
        out['depth_ndc'] = torch.empty(n, device=dev)
        out['depth_var_ndc'] = torch.empty(n, device=dev)
    out['alpha'] = torch.empty((n, s), device=dev)
    if want_weights:
        out['weights'] = torch.empty((n, s), device=dev)
    flags = (FLAG_NDC if ndc else 0) | (FLAG_WHITE_BKGD if white_bkgd else 0)
    nbytes = _lib.load().snerf_render_workspace_bytes(C.byref(desc), n, s, flags)
    if nbytes == 0:
        raise RuntimeError(f'snerf_render_workspace_bytes: {_lib.load().snerf_last_error().decode()}')
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    LAUNCHES['count'] += 3          # view-bias table, the MLP kernel with the compositing arithmetic in its epilogue, the per-ray fold
    with _timed('mlp_forward'):
        _lib.check(_lib.load().snerf_render_forward(
            C.byref(desc), pointer_table(params), _ptr(packed, torch.uint8), _ptr(pts_o), _ptr(pts_d), _ptr(view_dirs), _ptr(z),
            _ptr(rays_o), _ptr(rays_d), _ptr(out['rgb']), _ptr(out['acc']), _ptr(out['depth']), _ptr(out['depth_var']),
            _ptr(out.get('depth_ndc')), _ptr(out.get('depth_var_ndc')), _ptr(out['alpha']), _ptr(out.get('weights')),
            _ptr(ws, torch.uint8), ws.numel(), n, s, flags, _stream()), 'snerf_render_forward')
    return out


def mlp_backward(desc: MlpDesc, params, packed, rays_o, rays_d, view_dirs, z, sigma, rgb, d_sigma, d_rgb,
                 grads: List[Optional[torch.Tensor]], workspace, flags: int) -> None:
    n, s = z.shape
    LAUNCHES['count'] += (45 if flags & FLAG_PRECISE else 2)   # tensor path: dgrad chain + wgrad
    with _timed('mlp_backward') as tm:
        split = getattr(tm, 'split_event', None)        # bench.py: an event recorded between the dgrad and the wgrad launch
        lib = _lib.load()
        if split is not None:
            lib.snerf_set_backward_split_event(split.cuda_event)
        try:
            _lib.check(lib.snerf_mlp_backward(
                C.byref(desc), pointer_table(params), _ptr(packed, torch.uint8), _ptr(rays_o), _ptr(rays_d), _ptr(view_dirs),
                _ptr(z), _ptr(sigma), _ptr(rgb), _ptr(d_sigma), _ptr(d_rgb), pointer_table(grads), _ptr(workspace, torch.uint8),
                workspace.numel(), n, s, flags, _stream()), 'snerf_mlp_backward')
        finally:
            if split is not None:
                lib.snerf_set_backward_split_event(None)


def visibility_forward(desc: MlpDesc, params, mlp_workspace, rays_o, rays_d, z, rays_o2, flags: int):
    """Secondary-view visibility head on the workspace of the same MLP's `mlp_forward` (row a14 / N4; reference :317-325,
    :640-649, :687-715): precise path, or tensor path with FLAG_VIS_HEAD in the forward's flags -- pass those flags (+ FLAG_NDC).
    rays_o2 [N, nf-1, 3] or None.  -> visibility [N,S], visibility2 [N,S,nf-1] or None, workspace."""
    n, s = z.shape
    n_other = 0 if rays_o2 is None else rays_o2.shape[1]
    rays_o2 = None if rays_o2 is None else _f32(rays_o2)
    vis = torch.empty((n, s), device=z.device)
    vis2 = torch.empty((n, s, n_other), device=z.device) if n_other else None
    # the tensor path keeps everything in the MLP workspace; the precise path has buffers of its own
    nbytes = _lib.load().snerf_visibility_workspace_bytes(C.byref(desc), n, s, n_other) if flags & FLAG_PRECISE else 256
    ws = torch.empty(nbytes, dtype=torch.uint8, device=z.device)
    LAUNCHES['count'] += (3 + 3 * n_other) if flags & FLAG_PRECISE else 1
    _lib.check(_lib.load().snerf_visibility_forward(
        C.byref(desc), pointer_table(params), _ptr(mlp_workspace, torch.uint8), _ptr(rays_o), _ptr(rays_d), _ptr(z), _ptr(rays_o2),
        _ptr(vis), _ptr(vis2), _ptr(ws, torch.uint8), ws.numel(), n, s, n_other, flags, _stream()), 'snerf_visibility_forward')
    return vis, vis2, ws


def visibility_backward(desc: MlpDesc, params, mlp_workspace, rays_o, rays_d, z, rays_o2, vis, vis2, d_vis, d_vis2,
                        grads: List[Optional[torch.Tensor]], workspace, flags: int) -> None:
    """Run BEFORE mlp_backward(flags | FLAG_VIS_GRAD): adds the fourth-row / view-layer gradients to `grads` and leaves
    d hv and d feature in the MLP workspace."""
    n, s = z.shape
    n_other = 0 if rays_o2 is None else rays_o2.shape[1]
    rays_o2 = None if rays_o2 is None else _f32(rays_o2)
    LAUNCHES['count'] += (4 + 7 * n_other) if flags & FLAG_PRECISE else 1
    _lib.check(_lib.load().snerf_visibility_backward(
        C.byref(desc), pointer_table(params), _ptr(mlp_workspace, torch.uint8), _ptr(rays_o), _ptr(rays_d), _ptr(z), _ptr(rays_o2),
        _ptr(vis), _ptr(vis2), _ptr(None if d_vis is None else _f32(d_vis)), _ptr(None if d_vis2 is None else _f32(d_vis2)),
        pointer_table(grads), _ptr(workspace, torch.uint8), workspace.numel(), n, s, n_other, flags, _stream()),
        'snerf_visibility_backward')


def visibility2_composite_forward(weights, acc, vis2) -> torch.Tensor:
    """visibility2 map [N, nf-1] = sum_s w vis2 / (acc + 1e-6)  (reference :479-482)."""
    n, s, nv = vis2.shape
    out = torch.empty((n, nv), device=vis2.device)
    LAUNCHES['count'] += 1
    _lib.check(_lib.load().snerf_visibility2_composite_forward(_ptr(weights), _ptr(acc), _ptr(vis2), _ptr(out), n, s, nv, _stream()),
               'snerf_visibility2_composite_forward')
    return out


def visibility2_composite_backward(weights, acc, vis2, vis2_map, g_map):
    """-> d_vis2 [N,S,nf-1], d_weights [N,S], d_acc [N] (the last two join the compositing backward's incoming gradients)."""
    n, s, nv = vis2.shape
    d_vis2, d_w, d_acc = torch.empty_like(vis2), torch.empty((n, s), device=vis2.device), torch.empty(n, device=vis2.device)
    LAUNCHES['count'] += 1
    _lib.check(_lib.load().snerf_visibility2_composite_backward(_ptr(weights), _ptr(acc), _ptr(vis2), _ptr(vis2_map), _ptr(_f32(g_map)),
                                                                _ptr(d_vis2), _ptr(d_w), _ptr(d_acc), n, s, nv, _stream()),
               'snerf_visibility2_composite_backward')
    return d_vis2, d_w, d_acc


def tensor_selftest() -> List[float]:
    errs = (C.c_float * 4)()
    _lib.check(_lib.load().snerf_tensor_selftest(errs, _stream()), 'snerf_tensor_selftest')
    return list(errs)
