"""How much heavier is the reference's real criterion (pasco_b200/criterion.py: 4 matchings per subnet, voxel<->query
consistency terms on the 3 aux levels) than the compact loss the benchmark steps on?  CPU-only measurement on the fixed
1/16-scene crop of the CPU arm (oracle network), forward + backward of the LOSS alone.      python tools/criterion_cost_cpu.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

import net_oracle  # noqa: E402
from scene_and_loss import make_scene, total_loss  # noqa: E402
from pasco_b200 import criterion as CR  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    div = net_oracle.CROP_DIV
    net = net_oracle.OracleNet(caps=tuple(max(8, c // div) for c in (25000, 120000, 400000))).train()
    scene = make_scene(net_oracle.CROP_GRID, 0.10, 1, seed=0)
    freq = {f"1_{s}": np.ones(20) for s in (1, 2, 4)}
    t0 = time.time()
    out = net(scene)
    t_fwd = time.time() - t0
    X, Y, Z = net_oracle.CROP_GRID
    sem = scene["sem_labels"]["1_1"]
    masks = torch.zeros(len(scene["mask_boxes"]), X, Y, Z, dtype=torch.bool)
    for i, (lo, hi) in enumerate(scene["mask_boxes"]):
        masks[i, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = True
    batch = {"sem_labels": scene["sem_labels"], "min_Cs": scene["min_Cs"], "max_Cs": scene["max_Cs"], "semantic_label": sem,
             "mask_label": [{"labels": scene["mask_classes"], "masks": masks}],
             "geo_labels": {"1_1": (sem > 0).float()}}
    cw = torch.ones(21)
    cw[0] = cw[-1] = 0.1
    crit = CR.SetCriterion(20, [cw], torch.ones(20))
    leaves = [p for p in net.parameters()]

    def timed(fn):
        ts = []
        for _ in range(3):
            t0 = time.time()
            loss = fn()
            torch.autograd.grad(loss, [out["panop_predictions"][0]["voxel_logits"].F] +
                                [lg.F for per in out["sem_logits_at_scales"].values() for lg in per], retain_graph=True,
                                allow_unused=True)
            ts.append(time.time() - t0)
        return sorted(ts)[1], float(loss)
    t_compact, l_compact = timed(lambda: total_loss(out, scene, 20, freq))
    t_full, l_full = timed(lambda: CR.training_loss(out, batch, crit, freq)[0])
    n = out["panop_predictions"][0]["voxel_logits"].F.shape[0]
    print({"crop": net_oracle.CROP_GRID, "mask_rows": n, "network_forward_s": round(t_fwd, 2),
           "compact_loss_fwd_bwd_s": round(t_compact, 3), "reference_criterion_fwd_bwd_s": round(t_full, 3),
           "loss_values": [round(l_compact, 3), round(l_full, 3)], "threads": torch.get_num_threads()})
    del leaves


if __name__ == "__main__":
    main()
