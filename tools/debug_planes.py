"""Step-by-step check of the plane-gather kernels against the register-gather ones (prints before every launch)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pasco_b200 import ops  # noqa: E402


def say(*a):
    print(*a, flush=True)


def main():
    which = sys.argv[1:] or ["fwd", "dgrad", "wgrad"]
    for (shape, occ, cin, cout) in (((24, 20, 12), 0.3, 64, 64), ((48, 40, 14), 0.35, 64, 64), ((48, 40, 14), 0.35, 128, 128),
                                    ((48, 40, 14), 0.35, 256, 256), ((128, 128, 32), 0.4, 64, 64)):
        g = torch.Generator().manual_seed(4)
        occm = torch.rand(*shape, generator=g) < occ
        c = torch.nonzero(occm).int()
        C = torch.cat([torch.zeros(c.shape[0], 1, dtype=torch.int32), c], 1).cuda()
        N = C.shape[0]
        table, _ = ops.hash_insert(C)
        nbr = ops.kernel_map_probe(C, table, 3, (1, 1, 1))
        F = torch.randn(N, cin, generator=g).cuda()
        G = torch.randn(N, cout, generator=g).cuda()
        W = (torch.randn(27, cin, cout, generator=g) * 0.05).cuda()
        koff = [26 - k for k in range(27)]
        for prec in ("fp32", "bf16"):
            ops.set_precision(prec)
            ops.split_k(False)
            say(f"--- N={N} {cin}->{cout} {prec}")
            ops.use_planes(False)
            f0 = ops.conv_apply(F, W, nbr, N, False, None)
            d0 = ops.conv_apply(G, W, nbr, N, True, koff)
            w0 = ops.conv_wgrad(F, G, nbr, 27, cin, cout)
            torch.cuda.synchronize()
            ops.use_planes(True)
            say("split");
            fp, gp = ops.split_planes(F), ops.split_planes(G)
            torch.cuda.synchronize()
            if "fwd" in which:
                say("fwd planes")
                f1 = ops.conv_apply(F, W, nbr, N, False, None, planes=fp)
                torch.cuda.synchronize()
                say("  fwd equal:", torch.equal(f0, f1), float((f0 - f1).abs().max()))
            if "dgrad" in which:
                say("dgrad planes")
                d1 = ops.conv_apply(G, W, nbr, N, True, koff, planes=gp)
                torch.cuda.synchronize()
                say("  dgrad equal:", torch.equal(d0, d1), float((d0 - d1).abs().max()))
            if "wgrad" in which:
                say("wgrad planes")
                w1 = ops.conv_wgrad(F, G, nbr, 27, cin, cout, in_planes=fp, g_planes=gp)
                torch.cuda.synchronize()
                say("  wgrad rel:", float((w0 - w1).abs().max() / w0.abs().max()))
    ops.set_precision("fp32")
    say("done")


if __name__ == "__main__":
    main()
