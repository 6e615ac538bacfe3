"""Diagnostic: layer-by-layer forward/backward comparison of the mini U-Net slice (tests/test_gpu_ops.py)
between the CPU oracle and the CUDA engine."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import me_oracle as OR
from pasco_b200 import build
build.build()
from pasco_b200 import me as ME, ops
import test_gpu_ops as T

ops.set_precision("fp32")
torch.manual_seed(0)
rnet, gnet = T._mini_net(OR), T._mini_net(ME)
gnet.load_state_dict(rnet.state_dict())
gnet.cuda()
C, F = T.scene(shape=(20, 16, 8), p=0.25, C=16, batch=1)


def run(net, M, x):
    acts, grads = {}, {}
    hooks = []
    for name, mod in net.named_modules():
        if len(list(mod.children())) == 0 or name in ("r1", "r2", "r3"):
            def hk(m, i, o, name=name):
                if hasattr(o, "F"):
                    acts[name] = o
                    if o.F.requires_grad:
                        o.F.register_hook(lambda g, name=name: grads.__setitem__(name, g))
            hooks.append(mod.register_forward_hook(hk))
    y = net(x)
    (y.F ** 2).mean().backward()
    return acts, grads


ra, rg = run(rnet, OR, OR.SparseTensor(F, C))
ga, gg = run(gnet, ME, ME.SparseTensor(F.cuda(), C.cuda()))


def canon(st, t):
    Cc = st.C.detach().cpu()
    order = torch.argsort(OR.pack_keys(Cc))
    return t.detach().cpu().double()[order]


print(f"{'layer':28s} {'fwd maxnorm':>12s} {'fwd /std':>12s} {'grad maxnorm':>12s}")
for name in ra:
    if name not in ga:
        continue
    a, b = canon(ga[name], ga[name].F), canon(ra[name], ra[name].F)
    e = (a - b).abs()
    chstd = b.std(0).clamp(min=1e-12)
    line = f"{name:28s} {float(e.max() / b.abs().max()):12.2e} {float((e / chstd).max()):12.2e}"
    if name in rg and name in gg:
        ga_, gb_ = canon(ga[name], gg[name]), canon(ra[name], rg[name])
        line += f" {float((ga_ - gb_).abs().max() / gb_.abs().max()):12.2e}"
    print(line)
