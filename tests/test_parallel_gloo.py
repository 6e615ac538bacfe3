"""world_size-2 gloo tests (CPU) of the N>1 host logic: flat gradient all-reduce, packed SyncBN statistics,
disjoint scene sharding."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pasco_b200 import parallel
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    params = list(model.parameters())
    x = torch.randn(5, 8, generator=torch.Generator().manual_seed(100 + rank))
    model[:3](x).square().mean().backward()           # last layer unused → grad None (find_unused_parameters)
    local = [None if p.grad is None else p.grad.clone() for p in params]
    bucket = parallel.allreduce_gradients(params)
    gathered = [None] * world
    dist.all_gather_object(gathered, local)
    ok = True
    for i, p in enumerate(params):
        parts = [g[i] if g[i] is not None else torch.zeros_like(p) for g in gathered]
        ok &= torch.allclose(p.grad, sum(parts) / world, atol=1e-7)
    bucket2 = parallel.allreduce_gradients(params, bucket=bucket)
    ok &= bucket2 is bucket
    # bucketed reducer overlapped with backward: gradients are views into one flat buffer, unused parameters reduce as zeros
    model2 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    model2.load_state_dict(model.state_dict())
    params2 = list(model2.parameters())
    red = parallel.GradReducer(params2, bucket_mb=16 * 4 / (1 << 20))          # 16-float buckets → several buckets
    ok &= len(red.buckets) >= 3
    for _ in range(2):                                                          # two steps: zero_grad re-arms the hooks
        red.zero_grad()
        model2[:3](x).square().mean().backward()
        red.finish()
        for i, p in enumerate(params2):
            parts = [g[i] if g[i] is not None else torch.zeros_like(p) for g in gathered]
            ok &= torch.allclose(p.grad, sum(parts) / world, atol=1e-7)
            ok &= p.grad.data_ptr() >= red.flat.data_ptr() and p.grad.data_ptr() < red.flat.data_ptr() + red.flat.numel() * 4
    red.detach()
    # variable-length row all-gather (the sparse logit exchange of MIMO head sharding)
    rows = torch.arange((3 + 2 * rank) * 4, dtype=torch.float32).view(3 + 2 * rank, 4) + 100 * rank
    got = parallel.all_gather_rows(rows)
    ok &= len(got) == world and all(g.shape[0] == 3 + 2 * r for r, g in enumerate(got))
    ok &= all(torch.equal(g, torch.arange((3 + 2 * r) * 4, dtype=torch.float32).view(3 + 2 * r, 4) + 100 * r) for r, g in enumerate(got))
    ok &= parallel.all_gather_rows(torch.zeros(0 if rank == 0 else 2, 3))[0].shape == (0, 3)
    # packed SyncBN statistics
    xs = torch.randn(10 + 3 * rank, 6, dtype=torch.float64, generator=torch.Generator().manual_seed(7 + rank))
    stats = torch.stack([xs.sum(0), (xs * xs).sum(0)])
    gstats, gcount = parallel.sync_bn_statistics(stats, xs.shape[0])
    allx = torch.cat([torch.randn(10 + 3 * r, 6, dtype=torch.float64, generator=torch.Generator().manual_seed(7 + r))
                      for r in range(world)])
    ok &= gcount == allx.shape[0] and torch.allclose(gstats[0] / gcount, allx.mean(0)) \
        and torch.allclose(gstats[1] / gcount - (gstats[0] / gcount) ** 2, allx.var(0, unbiased=False))
    seeds = [None] * world
    dist.all_gather_object(seeds, parallel.scene_seeds(rank, world, 4))
    ok &= len(set(sum(seeds, []))) == 4 * world
    ok &= parallel.max_over_ranks(float(rank + 1), "cpu") == float(world)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_data_parallel_host_logic_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
