import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasco_b200 import ops
g = torch.Generator().manual_seed(4)
occ = torch.rand(128, 128, 32, generator=g) < 0.4
c = torch.nonzero(occ).int()
C = torch.cat([torch.zeros(c.shape[0], 1, dtype=torch.int32), c], 1).cuda()
N = C.shape[0]
table, _ = ops.hash_insert(C)
nbr = ops.kernel_map_probe(C, table, 3, (1, 1, 1))
F = torch.randn(N, 64, generator=g).cuda(); G = torch.randn(N, 64, generator=g).cuda()
W = (torch.randn(27, 64, 64, generator=g) * 0.05).cuda()
ops.set_precision("bf16"); ops.split_k(False)
for which in sys.argv[1:] or ["fwd", "wgrad"]:
    print("running", which, flush=True)
    if which == "fwd":
        f1 = ops.conv_apply(F, W, nbr, N, False, None)
    else:
        w1 = ops.conv_wgrad(F, G, nbr, 27, 64, 64)
    torch.cuda.synchronize()
    print("  done", which, flush=True)
