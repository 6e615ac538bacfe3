"""Multi-GPU plumbing of the data-parallel path: one process per GPU, scenes are the
independent units, no collective on the data path.  Collectives (NCCL over NVLink on the box, gloo in the
CPU tests): one flat gradient all-reduce per optimiser step (the reference's DDP, scripts/train.py:213-216)
and the packed SyncBatchNorm statistics (unet3d_sparse_v2.py:172-175 + Trainer(sync_batchnorm=True))."""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def scene_seeds(rank: int, world: int, n: int) -> List[int]:
    """Disjoint synthetic-scene seeds per rank (weak scaling: every rank gets n scenes)."""
    return [1000 * rank + i for i in range(n)]


def allreduce_gradients(params: Sequence[torch.nn.Parameter], group=None, bucket: Optional[torch.Tensor] = None):
    """Average gradients over the group with ONE flat all-reduce (parameters without a gradient contribute
    zeros = DDP's find_unused_parameters=True).  Returns the flat bucket so it can be reused next step."""
    world = dist.get_world_size(group)
    if world == 1:
        return bucket
    total = sum(p.numel() for p in params)
    ref = next(p for p in params)
    if bucket is None or bucket.numel() != total or bucket.device != ref.device:
        bucket = torch.empty(total, dtype=torch.float32, device=ref.device)
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            bucket[off:off + n].zero_()
        else:
            bucket[off:off + n].copy_(p.grad.reshape(-1))
        off += n
    dist.all_reduce(bucket, group=group)
    bucket /= world
    off = 0
    for p in params:
        n = p.numel()
        g = bucket[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return bucket


def sync_bn_statistics(stats: torch.Tensor, count: float, group=None):
    """Packed SyncBatchNorm reduction: stats float64 [2,C] (Σx, Σx²) and the local row count travel in one
    [2C+1] all-reduce; returns (global stats, global count)."""
    packed = torch.cat([stats.reshape(-1), torch.tensor([float(count)], dtype=stats.dtype, device=stats.device)])
    dist.all_reduce(packed, group=group)
    return packed[:-1].view_as(stats), float(packed[-1].item())


def enable_sync_batchnorm(model: torch.nn.Module, group=None) -> int:
    """Give every fused BatchNorm holder of pasco_b200.net3d the process group (ME.MinkowskiSyncBatchNorm
    .convert_sync_batchnorm equivalent for the engine-native model).  Returns the number of layers switched."""
    from .net3d import BNorm
    n = 0
    g = group if group is not None else dist.group.WORLD
    for m in model.modules():
        if isinstance(m, BNorm):
            m.group = g
            n += 1
    return n


def max_over_ranks(ms: float, device) -> float:
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
