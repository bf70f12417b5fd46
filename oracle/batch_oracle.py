"""TEST INFRASTRUCTURE ONLY: torch-CPU restatement of the tensor part of the reference's cached batch assembly.
Pinned by tests/golden/batch.npz (oracle/make_golden_batch.py runs the unmodified reference methods)."""
import torch

RAY = {'rays_o': 'rays_o', 'rays_d': 'rays_d', 'view_dirs': 'view_dirs', 'pixel_id': 'pixel_id', 'near': 'near_array', 'far': 'far_array',
       'rays_o_ndc': 'rays_o_ndc', 'rays_d_ndc': 'rays_d_ndc', 'near_ndc': 'near_array_ndc', 'far_ndc': 'far_array_ndc'}
SD = {'sparse_depth_values': 'depths', 'sparse_depth_errors': 'reprojection_errors', 'sparse_depth_values_ndc': 'depths_ndc'}


def _fill(table, indices, mask):
    """`x = -1 * ones(...)`; `x[mask] = table[indices[mask]]` (src/data_preprocessors/DataPreprocessor01.py:577-590)."""
    out = -1 * torch.ones((indices.shape[0],) + tuple(table.shape[1:]), dtype=table.dtype)
    out[mask] = table[indices[mask]]
    return out


def assemble_batch(nerf, sd, indices, mask_nerf, mask_sd=None):
    """load_nerf_cached_batch :572-620 followed by load_sparse_depth_cached_batch :655-700 (ndc=True)."""
    rows = mask_nerf if mask_sd is None else (mask_nerf | mask_sd)     # the second pass fills the sparse-depth rows of the ray tables
    batch = {k: _fill(nerf[src], indices, rows) for k, src in RAY.items()}
    batch['target_rgb'] = _fill(nerf['target_rgb'], indices, mask_nerf)
    if mask_sd is not None:
        batch.update({k: _fill(sd[src], indices, mask_sd) for k, src in SD.items()})
    return batch
