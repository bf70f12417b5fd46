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
