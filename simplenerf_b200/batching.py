"""Batch assembly on the device: host mirror of the tensor part of
`DataPreprocessor.load_cached_next_batch` (src/data_preprocessors/DataPreprocessor01.py:514-551).

`assemble_batch(tables, indices, mask_nerf, mask_sd, ndc)` returns the per-ray entries of the reference's batch dict
(same keys, shapes, dtypes and -1 fill) from the cached per-pixel tables, in ONE launch of `snerf_gather_rows` instead
of a fill, two boolean-mask index operations and a scatter per tensor.  `tables` holds the reference's
`preprocessed_data_dict['nerf_data']` tensors (rays_o, rays_d, view_dirs, pixel_id, target_rgb, near_array, far_array and
their *_ndc variants) and, when there are sparse-depth rays, `preprocessed_data_dict['sparse_depth_data']` under the key
'sparse_depth_data' (depths, reprojection_errors, depths_ndc).  There is no CPU path."""
from __future__ import annotations

import struct
from typing import Dict, Optional

import torch

from . import _lib, ops

_MINUS_ONE_F32 = struct.unpack('<I', struct.pack('<f', -1.0))[0]
_MINUS_ONE_I32 = 0xFFFFFFFF

# batch key -> nerf_data key; gathered on the NeRF rows and on the sparse-depth rows (:584-590 and :670-675)
_RAY_KEYS = {'rays_o': 'rays_o', 'rays_d': 'rays_d', 'view_dirs': 'view_dirs', 'pixel_id': 'pixel_id', 'near': 'near_array',
             'far': 'far_array'}
_RAY_KEYS_NDC = {'rays_o_ndc': 'rays_o_ndc', 'rays_d_ndc': 'rays_d_ndc', 'near_ndc': 'near_array_ndc', 'far_ndc': 'far_array_ndc'}
# batch key -> sparse_depth_data key; sparse-depth rows only (:679-684, :697-699)
_SD_KEYS = {'sparse_depth_values': 'depths', 'sparse_depth_errors': 'reprojection_errors'}
_SD_KEYS_NDC = {'sparse_depth_values_ndc': 'depths_ndc'}


def _mask8(m: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if m is None:
        return None
    return m.contiguous().view(torch.uint8) if m.dtype == torch.bool else m.to(torch.uint8).contiguous()


def gather_rows(entries, indices: torch.Tensor, check_bounds: bool = False) -> None:
    """entries: list of (src [M, ...], dst [N, ...], mask [N] or None); dst[i] = src[indices[i]] where the mask holds, else -1.
    check_bounds: validate 0 <= indices < M for every table first (one device synchronisation; the kernel itself does not check)."""
    n = indices.shape[0]
    if not indices.is_cuda:
        raise RuntimeError('simplenerf_b200 kernels need CUDA tensors (no CPU fallback exists)')
    indices = indices.to(torch.int64).contiguous()
    if check_bounds and n:
        lo, hi = int(indices.min()), int(indices.max())
        rows = min(src.shape[0] for src, _, _ in entries)
        if lo < 0 or hi >= rows:
            raise IndexError(f'gather_rows: indices span [{lo}, {hi}] but the smallest source table has {rows} rows')
    lib = _lib.load()
    for i in range(0, len(entries), _lib.GATHER_MAX_TABLES):
        part = entries[i:i + _lib.GATHER_MAX_TABLES]
        table = (_lib.GatherTable * len(part))()
        keep = []
        for k, (src, dst, mask) in enumerate(part):
            if src.dtype != dst.dtype or src.shape[1:] != dst.shape[1:] or dst.shape[0] != n or src.element_size() != 4:
                raise RuntimeError(f'table {i + k}: source {tuple(src.shape)} {src.dtype} vs destination {tuple(dst.shape)} {dst.dtype}')
            src = src.contiguous()
            m8 = _mask8(mask)
            keep.append((src, m8))
            table[k].src, table[k].dst = src.data_ptr(), ops._ptr(dst, dst.dtype)
            table[k].mask = None if m8 is None else ops._ptr(m8, torch.uint8)
            table[k].row_bytes = 4 * (src[0].numel() if src.dim() > 1 else 1)
            table[k].fill_bits = _MINUS_ONE_F32 if dst.dtype == torch.float32 else _MINUS_ONE_I32
        ops.LAUNCHES['count'] += 1
        _lib.check(lib.snerf_gather_rows(table, len(part), ops._ptr(indices, torch.int64), n, ops._stream()), 'snerf_gather_rows')


def assemble_batch(tables: Dict, indices: torch.Tensor, mask_nerf: torch.Tensor, mask_sd: Optional[torch.Tensor] = None,
                   ndc: bool = True) -> Dict[str, torch.Tensor]:
    n, dev = indices.shape[0], indices.device
    both = None if mask_sd is None else (mask_nerf | mask_sd)       # rows that end up filled after both reference passes
    ray_mask = mask_nerf if mask_sd is None else both
    out: Dict[str, torch.Tensor] = {}
    entries = []

    def add(key, src, mask):
        out[key] = torch.empty((n,) + tuple(src.shape[1:]), device=dev, dtype=src.dtype)
        entries.append((src, out[key], mask))

    for key, src_key in {**_RAY_KEYS, **(_RAY_KEYS_NDC if ndc else {})}.items():
        add(key, tables[src_key], ray_mask)
    add('target_rgb', tables['target_rgb'], mask_nerf)                                           # :588, NeRF rows only
    if mask_sd is not None:
        sd = tables['sparse_depth_data']
        for key, src_key in {**_SD_KEYS, **(_SD_KEYS_NDC if ndc else {})}.items():
            add(key, sd[src_key], mask_sd)
    gather_rows(entries, indices)
    return out


class HostBatchStager:
    """Host -> device input pipeline for batches that are assembled on the host (the reference's loader
    hands the model a dict of ~12 small tensors): ONE pinned staging buffer per slot, ONE host -> device copy per batch on a
    copy stream, double buffered, so the copy of batch i+1 travels under the kernels of batch i.

        stager = HostBatchStager(example_batch, device)          # layout from an example (keys, shapes, dtypes)
        stager.wait_host(slot); host = stager.host_views(slot)    # dict of pinned views: fill them in place (or `stage` copies)
        stager.submit(slot)                                       # async copy on the copy stream
        batch = stager.device_batch(slot)                         # the compute stream waits for that copy only
        ...                                                       # run the step on `batch`
        stager.release(slot)                                      # the slot may be overwritten once these kernels are done
    Non-tensor entries of the example (iter_num, num_frames, common_data) are passed through by reference."""

    def __init__(self, example: Dict, device, slots: int = 2):
        self.device = torch.device(device)
        self.layout = []                      # (key, offset, nbytes, shape, dtype)
        self.passthrough = {k: v for k, v in example.items() if not isinstance(v, torch.Tensor)}
        off = 0
        for k, v in example.items():
            if isinstance(v, torch.Tensor):
                nbytes = v.numel() * v.element_size()
                self.layout.append((k, off, nbytes, tuple(v.shape), v.dtype))
                off += (nbytes + 255) // 256 * 256
        self.nbytes = max(off, 256)
        self.host = [torch.empty(self.nbytes, dtype=torch.uint8).pin_memory() for _ in range(slots)]
        self.dev = [torch.empty(self.nbytes, dtype=torch.uint8, device=self.device) for _ in range(slots)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.copied = [torch.cuda.Event() for _ in range(slots)]
        self.submitted = [False] * slots      # a host -> device copy out of the slot's pinned buffer has been enqueued
        self.free = [None] * slots            # event after the last kernel that read the slot
        self.handed_out = [False] * slots     # device_batch(slot) was called and release(slot) has not been yet

    def _views(self, buf: torch.Tensor) -> Dict:
        out = dict(self.passthrough)
        for k, off, nbytes, shape, dtype in self.layout:
            out[k] = buf[off:off + nbytes].view(dtype).view(shape)
        return out

    def wait_host(self, slot: int) -> None:
        """Block the host until the last copy OUT of the slot's pinned buffer has finished: the host runs ahead of the device,
        and overwriting the buffer earlier would change the data of a copy that is still queued."""
        if self.submitted[slot]:
            self.copied[slot].synchronize()

    def host_views(self, slot: int) -> Dict:
        """Pinned views to fill in place; call `wait_host(slot)` first when the slot has been submitted before."""
        return self._views(self.host[slot])

    def stage(self, slot: int, batch: Dict) -> None:
        """Copy a host batch into the slot's pinned buffer (skip when the loader fills `host_views` directly) and submit."""
        self.wait_host(slot)
        views = self.host_views(slot)
        for k, _, _, _, _ in self.layout:
            views[k].copy_(batch[k])
        self.submit(slot)

    def submit(self, slot: int) -> None:
        if self.handed_out[slot]:
            # the caller forgot release(slot): everything it has enqueued so far on the compute stream may still read the
            # slot's device buffer, so order the copy after all of it (conservative) instead of racing
            self.release(slot)
        with torch.cuda.stream(self.copy_stream):
            if self.free[slot] is not None:
                self.copy_stream.wait_event(self.free[slot])
            self.dev[slot].copy_(self.host[slot], non_blocking=True)
            self.copied[slot].record(self.copy_stream)
        self.submitted[slot] = True

    def device_batch(self, slot: int) -> Dict:
        if not self.submitted[slot]:
            raise RuntimeError(f'HostBatchStager.device_batch({slot}): nothing has been submitted to this slot')
        torch.cuda.current_stream(self.device).wait_event(self.copied[slot])
        self.handed_out[slot] = True
        return self._views(self.dev[slot])

    def release(self, slot: int) -> None:
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.free[slot] = ev
        self.handed_out[slot] = False

    @property
    def bytes_per_batch(self) -> int:
        return sum(n for _, _, n, _, _ in self.layout)
