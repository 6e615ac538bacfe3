import torch


def dim_zero_cat(x):
    if isinstance(x, torch.Tensor):
        return x
    x = [y.unsqueeze(0) if y.ndim == 0 else y for y in x]
    return torch.cat(x, dim=0) if x else torch.zeros(0)
