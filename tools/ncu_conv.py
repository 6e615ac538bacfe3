"""Few launches of the conv kernels at the benchmark's largest layer shape (for ncu).
    ncu --set full -k regex:k_conv_tc -c 2 -o gpurun_out/conv python tools/ncu_conv.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasco_b200 import build
build.build()
from pasco_b200 import ops

occ = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
ch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
o = torch.rand(256, 256, 32, generator=g) < occ
c = torch.nonzero(o).int()
C = torch.cat([torch.zeros(c.shape[0], 1, dtype=torch.int32), c], 1).to(dev)
N = C.shape[0]
table, _ = ops.hash_insert(C)
nbr = ops.kernel_map_probe(C, table, 3, (1, 1, 1))
F = torch.randn(N, ch, device=dev)
W = torch.randn(27, ch, ch, device=dev) * 0.05
G = torch.randn(N, ch, device=dev)
pk = ops.PackedWeights()
ops.set_precision("fp32")
for _ in range(2):
    ops.conv_apply(F, W, nbr, N, False, None, packs=pk)
    ops.conv_wgrad(F, G, nbr, 27, ch, ch)
torch.cuda.synchronize()
print("done", N, int((nbr >= 0).sum()))
