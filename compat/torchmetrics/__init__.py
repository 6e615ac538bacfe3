"""Subset of torchmetrics used by pasco/models/metrics.py:7-10 (states + compute only)."""
import torch


class Metric(torch.nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self._defaults = {}

    def add_state(self, name, default, dist_reduce_fx=None):
        self._defaults[name] = default
        setattr(self, name, [] if isinstance(default, list) else default.clone())

    def reset(self):
        for n, d in self._defaults.items():
            setattr(self, n, [] if isinstance(d, list) else d.clone())

    def __call__(self, *a, **k):
        return self.update(*a, **k)
