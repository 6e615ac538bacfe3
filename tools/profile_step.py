"""One training step of the benchmark workload under torch.profiler: top CUDA kernels + wall-clock split.
    python tools/profile_step.py [--out gpurun_out/profile_step.txt]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasco_b200 import build  # noqa: E402
build.build()
from pasco_b200 import ops  # noqa: E402
from pasco_b200.net3d import PascoNet  # noqa: E402
from pasco_b200.losses import total_loss  # noqa: E402
from pasco_b200.synthetic import make_scene  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/profile_step.txt")
ap.add_argument("--precision", default="fp32")
a = ap.parse_args()
dev = torch.device("cuda")
ops.set_precision(a.precision)
torch.manual_seed(0)
net = PascoNet().to(dev).train()
opt = torch.optim.AdamW(net.parameters(), lr=1e-4, fused=True)
sc = bench.to_device(make_scene(seed=0), dev)


def sync_t():
    torch.cuda.synchronize()
    return time.time()


def step(profile_phases=False):
    t0 = sync_t()
    out = net(sc["in_feats"], sc["in_coords"], sc["global_min_Cs"], sc["global_max_Cs"], sc["min_Cs"], sc["max_Cs"])
    t1 = sync_t()
    loss = total_loss(out, sc, 20, net.class_frequencies)
    t2 = sync_t()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    t3 = sync_t()
    opt.step()
    t4 = sync_t()
    return dict(fwd=t1 - t0, loss=t2 - t1, bwd=t3 - t2, opt=t4 - t3)


for _ in range(2):
    step()
ph = step()
lines = ["phase seconds: " + ", ".join(f"{k}={v * 1e3:.1f}ms" for k, v in ph.items())]
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
lines.append(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
lines.append(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=70))
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
open(a.out, "w").write("\n".join(lines))
print(lines[0])
