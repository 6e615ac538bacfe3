"""MIMO head sharding check + timing (BASELINE configs[3]: M=3 subnets one per GPU, NCCL all-gather of the sparse logits).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 \
        tools/mimo_shard_check.py [--grid 256 256 32] [--steps 5]

Every rank runs the shared trunk; rank r < 3 runs subnet r's panoptic heads; results are all-gathered and compared with the
un-sharded forward of the same network on rank 0 (same weights, same scene), then both paths are timed (eval forward,
CUDA events, max over ranks) and the ensemble (pasco_b200.ensemble) runs on the gathered rows.  One JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, nargs=3, default=[256, 256, 32])
    ap.add_argument("--occ", type=float, default=0.10)
    ap.add_argument("--m", type=int, default=3)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    from pasco_b200 import build
    build.build()
    from pasco_b200 import ops, parallel, ensemble
    from pasco_b200.net3d import PascoNet
    from pasco_b200.synthetic import make_scene
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ops.set_precision("fp32")
    torch.manual_seed(0)
    net = PascoNet(n_classes=20, n_infers=a.m, in_channels=283, f=64, num_queries=100).to(dev).train()
    sc = make_scene(a.grid, a.occ, a.m, seed=0)
    d = lambda t: t.to(dev)  # noqa: E731
    scd = dict(sc, in_feats=[d(t) for t in sc["in_feats"]], in_coords=[d(t) for t in sc["in_coords"]])

    def full():
        with torch.no_grad():
            return net(scd["in_feats"], scd["in_coords"], sc["global_min_Cs"], sc["global_max_Cs"], sc["min_Cs"], sc["max_Cs"], test=True)

    def sharded():
        with torch.no_grad():
            return parallel.sharded_mimo_forward(net, scd, test=True)

    def timed(fn):
        for _ in range(2):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), out

    ms_full, ref = timed(full)
    ms_shard, got = timed(sharded)
    if rank == 0:
        diffs = {}
        for m in range(a.m):
            r, g = ref["panop_predictions"][m], got["panop_predictions"][m]
            same_rows = r["voxel_logits"].F.shape == g["voxel_logits"][0].shape and torch.equal(r["voxel_logits"].C, g["voxel_logits"][1])
            dq = float((r["query_logits"] - g["query_logits"]).abs().max() / r["query_logits"].abs().max())
            dv = float((r["voxel_logits"].F - g["voxel_logits"][0]).abs().max() / r["voxel_logits"].F.abs().max()) if same_rows else None
            diffs[f"subnet{m}"] = {"rows": int(g["voxel_logits"][0].shape[0]), "same_rows": bool(same_rows), "query_rel": dq, "voxel_rel": dv}
        Ts = [torch.eye(4)] * a.m
        sem = [(f, c[:, 1:]) for f, c in got["sem_logits_pruneds"]]
        sem_d = ensemble.ensemble_sem_compl(sem, Ts)
        ens = ensemble.ensemble_panop([{"voxel_logits": (p["voxel_logits"][0], p["voxel_logits"][1][:, 1:]), "query_logits": p["query_logits"]}
                                       for p in got["panop_predictions"]], sem_d, Ts)
        print(json.dumps({"what": "MIMO head sharding", "M": a.m, "n_gpus": world, "grid": a.grid,
                          "ms_unsharded_1gpu_heads": round(ms_full, 2), "ms_sharded": round(ms_shard, 2),
                          "speedup": round(ms_full / ms_shard, 3), "parity_vs_unsharded": diffs,
                          "ensemble_queries_kept": int(ens[-1]["query_probs"].shape[1]), "ensemble_rows": int(ens[-1]["voxel_probs"][0].shape[0]),
                          "note": "padded rows of shorter subnets take part in attention in the un-sharded batch (reference quirk), "
                                  "so only the longest subnet is expected to match bit for bit"}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
