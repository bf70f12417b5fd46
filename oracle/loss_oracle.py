"""TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's cpu_baseline): torch-CPU restatement of the reference's masked
per-ray losses.  Pinned by tests/golden/losses.npz, generated from the unmodified reference modules by
oracle/make_golden_losses.py."""
import torch


def masked_mse(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """MSE01.compute_mse (src/loss_functions/MSE01.py:53-59) for [N,C] and SparseDepthMSE01.compute_depth_loss
    (SparseDepthMSE01.py:58-63) for [N]: mean over channels, then over the masked rays; 0 when no ray is masked in."""
    pred, target = pred[mask], target[mask]
    if pred.numel() == 0:
        return torch.zeros((), dtype=pred.dtype)
    sq = torch.square(pred - target)
    return torch.mean(torch.mean(sq, dim=1)) if sq.dim() == 2 else torch.mean(sq)


def total_loss(streams) -> torch.Tensor:
    """LossComputer.compute_losses (LossComputer01.py:40-50): sum of weight * loss over (pred, target, mask, weight)."""
    return sum(w * masked_mse(p, t, m) for p, t, m, w in streams)


# ------------------------------------------------------------------------------------------------
# patch-reprojection depth losses (PointsAugmentationDepthLoss02 / ViewsAugmentationDepthLoss02 /
# CoarseFineConsistencyLoss02: the three modules share compute_loss_nerf line for line)
# ------------------------------------------------------------------------------------------------
def closest_views(poses: torch.Tensor) -> torch.Tensor:
    """For every view the index of the nearest other view (PointsAugmentationDepthLoss02.py:126-130: the second
    smallest camera distance; the reference evaluates it per ray, it only depends on the ray's view)."""
    origins = poses[:, :3, 3]
    dist = torch.sqrt(torch.sum(torch.square(origins[:, None, :] - origins[None, :, :]), dim=2))
    return torch.kthvalue(dist, 2, dim=1)[1]


def projection_matrices(poses: torch.Tensor, intrinsics: torch.Tensor) -> torch.Tensor:
    """CommonUtils01.reproject :63-69 without the point: K[0] @ diag(1,-1,-1) @ R_b^T per view, evaluated left to right."""
    permuter = torch.eye(3)
    permuter[1:] *= -1
    return intrinsics[:1] @ permuter[None] @ poses[:, :3, :3].transpose(1, 2)


def reprojection_masks(depth1, depth2, mask_nerf, rays_o, rays_d, poses, images, pixel_ids, intrinsics, resolution,
                       patch=(5, 5), rmse_threshold=0.1):
    """compute_loss_nerf :98-165 up to the two masks, over the rays with mask_nerf.  -> mask1 (model 1 is the more
    accurate one: it supervises depth2), mask2, both [n_masked] bool."""
    h, w = resolution
    px, py = patch
    hpx, hpy = px // 2, py // 2
    pixel_ids = pixel_ids.long()
    image_ids = pixel_ids[:, 0]
    o, d = rays_o[mask_nerf], rays_d[mask_nerf]
    d1, d2 = depth1.detach()[mask_nerf], depth2.detach()[mask_nerf]
    ids_a, pix_a = image_ids[mask_nerf], pixel_ids[mask_nerf]
    ids_b = closest_views(poses)[image_ids][mask_nerf]
    proj = projection_matrices(poses, intrinsics)[ids_b]
    origins_b = poses[ids_b][:, :3, 3]

    def reproject(points):
        p = (proj @ (points - origins_b)[..., None])[..., 0]
        return (p[:, :2] / p[:, 2:]).round().long()                                             # :139-140

    pos1, pos2 = reproject(o + d * d1[:, None]), reproject(o + d * d2[:, None])                  # :136-137
    xa, ya = pix_a[:, 1], pix_a[:, 2]

    def valid(x, y):
        return (x >= hpx) & (x < w - hpx) & (y >= hpy) & (y < h - hpy)                           # :147-149

    va, v1, v2 = valid(xa, ya), valid(pos1[:, 0], pos1[:, 1]), valid(pos2[:, 0], pos2[:, 1])
    padded = torch.nn.functional.pad(images, (0, 0, 0, hpy, 0, hpx))                            # :156

    def patches(ids, x, y):
        x, y = torch.clip(x, 0, w - 1), torch.clip(y, 0, h - 1)                                  # :151-152
        rows = []
        for dy in range(-hpy, hpy + 1):
            rows.append(torch.stack([padded[ids, y + dy, x + dx] for dx in range(-hpx, hpx + 1)], 1))
        return torch.stack(rows, 1)                                                              # (n, py, px, 3)

    pa = patches(ids_a, xa, ya)
    rmse1 = torch.sqrt(torch.mean(torch.square(pa - patches(ids_b, pos1[:, 0], pos1[:, 1])), dim=(1, 2, 3)))   # :180
    rmse2 = torch.sqrt(torch.mean(torch.square(pa - patches(ids_b, pos2[:, 0], pos2[:, 1])), dim=(1, 2, 3)))
    mask1 = ((rmse1 < rmse2) | ~v2) & (rmse1 < rmse_threshold) & v1 & va                         # :167
    mask2 = ((rmse2 < rmse1) | ~v1) & (rmse2 < rmse_threshold) & v2 & va                         # :169
    return mask1, mask2


def reprojection_depth_loss(depth1, depth2, mask_nerf, rays_o, rays_d, poses, images, pixel_ids, intrinsics, resolution,
                            patch=(5, 5), rmse_threshold=0.1, symmetric=False):
    """compute_loss_nerf :98-176.  As written (:172-175) depth1 is pulled to depth2 where model 2 is the more accurate one
    (mask2) and vice versa, each term a mean over ALL rays with mask_nerf (compute_depth_mse :196-212).  As it RUNS, only
    the first term exists: compute_depth_mse zeroes pred_depth[~mask] and gt_depth[~mask] in place, and gt_depth is
    depth.detach(), an alias of the other call's pred_depth.  After the first call both depth vectors are zero outside
    mask2; mask1 and mask2 exclude each other, so the second call computes (0 - 0)^2 everywhere -- value 0, gradient 0
    (the fixtures generated from the unmodified modules show gradients for the main depth only).  symmetric=True gives
    the form the code reads as."""
    mask1, mask2 = reprojection_masks(depth1, depth2, mask_nerf, rays_o, rays_d, poses, images, pixel_ids, intrinsics,
                                      resolution, patch, rmse_threshold)
    d1, d2 = depth1[mask_nerf], depth2[mask_nerf]
    if d1.numel() == 0:
        return torch.zeros(())
    loss = torch.mean(torch.square(torch.where(mask2, d1 - d2.detach(), torch.zeros_like(d1))))
    if symmetric:
        loss = loss + torch.mean(torch.square(torch.where(mask1, d2 - d1.detach(), torch.zeros_like(d1))))
    return loss
