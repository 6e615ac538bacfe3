"""Diagnostic: MaskTransformer on CUDA (sparse attention mask through the engine) vs the same module on CPU
with the attention mask computed by the oracle's max-pool + lookup."""
import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import me_oracle as OR
from pasco_b200 import build
build.build()
from pasco_b200 import me as ME, ops
from pasco_b200.net3d import MaskTransformer

torch.manual_seed(0)
cpu = MaskTransformer([256, 128, 64], 20, 384, 100, 8, 1024, 64, 1)
gpu = copy.deepcopy(cpu).cuda()
g = torch.Generator().manual_seed(1)


def mk(scale, C, p):
    occ = torch.rand(32 // scale, 32 // scale, 8 // scale, generator=g) < p
    c = torch.nonzero(occ).int() * scale
    bc = OR.utils.batched_coordinates([c])
    return bc, torch.randn(bc.shape[0], C, generator=g)


feats = {4: mk(4, 256, 0.6), 2: mk(2, 128, 0.5), 1: mk(1, 64, 0.3)}


class FakeST:
    def __init__(s, F, C):
        s.F, s.C = F, C


cpu_masks = []


def attn_mask_cpu(self, mask_logits, vox1, src, scale):
    keep = OR.SparseTensor((mask_logits.detach() > 0).float(), vox1.C)
    if scale != 1:
        keep = OR.MinkowskiMaxPooling(kernel_size=scale, stride=scale, dimension=3)(keep)
    rows = OR.lookup(keep.C, src.C)
    at = torch.where(rows[:, None] >= 0, keep.F[rows.clamp(min=0)], torch.zeros(1))
    m = (at == 0).t().contiguous()
    m[m.all(1)] = False
    cpu_masks.append(m)
    return m


gpu_masks = []
orig = MaskTransformer.attn_mask


def attn_mask_gpu(self, mask_logits, vox1, src, scale):
    m = orig(self, mask_logits, vox1, src, scale)
    gpu_masks.append(m.cpu())
    return m


with torch.no_grad():
    MaskTransformer.attn_mask = attn_mask_cpu
    c_cls, c_msk = cpu({s: FakeST(F, C) for s, (C, F) in feats.items()})
    MaskTransformer.attn_mask = attn_mask_gpu
    mgr = ME.CoordinateManager()
    gf = {}
    for s, (C, F) in feats.items():
        gf[s] = ME.SparseTensor(F.cuda(), C.cuda(), tensor_stride=s, coordinate_manager=mgr)
    g_cls, g_msk = gpu(gf)
for i in range(len(c_msk)):
    print(i, "mask logits err", float((g_msk[i].cpu() - c_msk[i]).abs().max()), "of", float(c_msk[i].abs().max()),
          "class err", float((g_cls[i].cpu() - c_cls[i]).abs().max()))
for i, (a, b) in enumerate(zip(gpu_masks, cpu_masks)):
    print("attn mask", i, a.shape, b.shape, "mismatches", int((a != b).sum()), "masked frac", float(b.float().mean()))
