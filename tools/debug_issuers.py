"""Which rows / tiles differ between k_conv_pl (multi-issuer) and k_conv_tc?  PASCO_PL_DEPTH=n sets the issuer count."""
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
F = torch.randn(N, 64, generator=g).cuda()
W = (torch.randn(27, 64, 64, generator=g) * 0.05).cuda()
ops.set_precision("fp32"); ops.split_k(False)
ops.use_planes(False); f0 = ops.conv_apply(F, W, nbr, N, False, None); torch.cuda.synchronize()
ops.use_planes(True); f1 = ops.conv_apply(F, W, nbr, N, False, None); torch.cuda.synchronize()
G = torch.randn(N, 64, generator=g).cuda()
ops.use_planes(False); w0 = ops.conv_wgrad(F, G, nbr, 27, 64, 64); torch.cuda.synchronize()
ops.use_planes(True); w1 = ops.conv_wgrad(F, G, nbr, 27, 64, 64); torch.cuda.synchronize()
print("wgrad rel", float((w0 - w1).abs().max() / w0.abs().max()), "env", {k: v for k, v in os.environ.items() if k.startswith("PASCO_")}, flush=True)
bad = (f0 != f1).any(1)
print("issuers", os.environ.get("PASCO_PL_DEPTH"), "N", N, "bad rows", int(bad.sum()), "max", float((f0 - f1).abs().max()), flush=True)
if bad.any():
    rows = torch.nonzero(bad).flatten()
    tiles = rows // 128
    print("tiles mod 4 histogram", torch.bincount(tiles % 4, minlength=4).tolist())
    print("first bad rows", rows[:10].tolist(), "bad tiles", int(torch.unique(tiles).numel()), "of", (N + 127) // 128)
    r = int(rows[0]); print("cols bad in first row", torch.nonzero(f0[r] != f1[r]).flatten().tolist()[:20])
    grp = torch.unique(tiles // 4)
    print("bad groups", grp[:20].tolist(), "count", grp.numel(), "groups mod 148:", torch.unique(grp % 148)[:20].tolist())
