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
lines.append(prof.key_averages().table(sort_by="cpu_time_total", row_limit=40, max_name_column_width=70))
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
busy = sum(e.time_range.end - e.time_range.start for e in ev) / 1e3
span = (max(e.time_range.end for e in ev) - min(e.time_range.start for e in ev)) / 1e3
lines.insert(1, f"GPU busy {busy:.1f} ms of a {span:.1f} ms span ({len(ev)} kernels/copies)")
# GPU idle gaps: which CPU-side op was running while the GPU had nothing to do
ks = sorted(((e.time_range.start, e.time_range.end) for e in ev))
cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and e.time_range.end - e.time_range.start > 0]
cpu.sort(key=lambda e: e.time_range.start)
import bisect, collections
starts = [e.time_range.start for e in cpu]
gaps = collections.Counter()
gap_n = collections.Counter()
end = ks[0][1]
total_gap = 0.0
for st, en in ks[1:]:
    if st - end > 5:        # µs
        g = st - end
        total_gap += g
        mid = end + g / 2
        i = bisect.bisect_right(starts, mid)
        # outermost-but-informative: the LONGEST enclosing op that is not a profiler/step wrapper, and the innermost
        encl = [c for c in cpu[max(0, i - 400):i] if c.time_range.start <= mid <= c.time_range.end]
        if encl:
            encl.sort(key=lambda c: c.time_range.end - c.time_range.start)
            inner = encl[0].name
            outer = next((c.name for c in reversed(encl) if not c.name.startswith("ProfilerStep")), inner)
            key = f"{outer[:48]} > {inner[:40]}"
        else:
            key = "(no CPU op: python between ops)"
        gaps[key] += g
        gap_n[key] += 1
    end = max(end, en)
lines.insert(2, f"GPU idle (gaps > 5 us): {total_gap / 1e3:.1f} ms; by enclosing CPU op (outermost > innermost):\n" + "\n".join(
    f"   {v / 1e3:7.2f} ms  {gap_n[k]:4d}x  {k}" for k, v in gaps.most_common(40)))
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
open(a.out, "w").write("\n".join(lines))
print(lines[0])
