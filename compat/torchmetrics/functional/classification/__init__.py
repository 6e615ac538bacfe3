import torch


def binary_calibration_error(preds, target, n_bins=15, norm="l1", ignore_index=None, validate_args=True):
    """Expected calibration error for binary predictions (same binning as torchmetrics ≥0.11)."""
    preds = preds.float().flatten()
    target = target.flatten()
    if ignore_index is not None:
        keep = target != ignore_index
        preds, target = preds[keep], target[keep]
    if preds.numel() and ((preds < 0).any() or (preds > 1).any()):
        preds = preds.sigmoid()
    conf = torch.where(preds >= 0.5, preds, 1 - preds)
    acc = ((preds >= 0.5).long() == target.long()).float()
    bins = torch.linspace(0, 1, n_bins + 1, device=preds.device)
    idx = torch.bucketize(conf, bins, right=True) - 1
    idx = idx.clamp(0, n_bins - 1)
    cnt = torch.zeros(n_bins, device=preds.device).index_add_(0, idx, torch.ones_like(conf))
    sconf = torch.zeros(n_bins, device=preds.device).index_add_(0, idx, conf)
    sacc = torch.zeros(n_bins, device=preds.device).index_add_(0, idx, acc)
    nz = cnt > 0
    gap = (sacc[nz] / cnt[nz] - sconf[nz] / cnt[nz]).abs()
    prop = cnt[nz] / cnt.sum().clamp(min=1)
    if norm == "l1":
        return (gap * prop).sum()
    if norm == "max":
        return gap.max() if gap.numel() else torch.tensor(0.0)
    return torch.sqrt((gap ** 2 * prop).sum())
