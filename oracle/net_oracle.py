"""CPU ORACLE — test / baseline infrastructure, NOT product code.

PaSCo's hot path (voxeliser → Net3D sparse U-Net → MaskPLS transformer) composed from the oracle's
MinkowskiEngine restatement (oracle/me_oracle), module for module as the reference composes the real
MinkowskiEngine — unfused, one op per module, exactly the structure of:
  pasco/maskpls/mink.py:505-534,618-658   pasco/models/encoder_v2.py:89-183
  pasco/models/decoder_v3.py:77-172,319-511   pasco/models/unet3d_sparse_v2.py:15-86,182-256
  pasco/models/layers.py:646-726   pasco/models/transformer/transformer_predictor_v2.py:111-303
It exists because /root/reference cannot travel to the GPU box: bench.py's `cpu_baseline` and
`--impl reference` legs time THIS on the host cores ("CPU restatement of the MinkowskiEngine algorithm",
BASELINE.md §3).  Only bench.py / tests / smoke may import it.
"""
from __future__ import annotations

import math
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import me_oracle as ME  # noqa: E402


def conv_block(cin, cout, ks, stride):            # mink.py:505-518
    return nn.Sequential(ME.MinkowskiConvolution(cin, cout, kernel_size=ks, stride=stride, dimension=3),
                         ME.MinkowskiBatchNorm(cout), ME.MinkowskiLeakyReLU(inplace=True))


def deconv_block(cin, cout):                      # mink.py:520-534
    return nn.Sequential(ME.MinkowskiConvolutionTranspose(cin, cout, kernel_size=2, stride=2, dimension=3,
                                                          expand_coordinates=True),
                         ME.MinkowskiBatchNorm(cout), ME.MinkowskiLeakyReLU(inplace=True))


class ResidualBlock(nn.Module):                   # mink.py:618-658
    def __init__(self, c):
        super().__init__()
        self.net = nn.Sequential(ME.MinkowskiBatchNorm(c), ME.MinkowskiReLU(inplace=True),
                                 ME.MinkowskiConvolution(c, c, kernel_size=3, dimension=3),
                                 ME.MinkowskiBatchNorm(c), ME.MinkowskiReLU(inplace=True),
                                 ME.MinkowskiConvolution(c, c, kernel_size=3, dimension=3))
        self.relu = ME.MinkowskiReLU(inplace=True)

    def forward(self, x):
        return self.relu(x + self.net(x))


def down_stage(cin, cout):                        # encoder_v2.py:122-130
    return nn.Sequential(conv_block(cin, cout, 2, 2), ME.MinkowskiBatchNorm(cout), ME.MinkowskiReLU(),
                         ResidualBlock(cout), ResidualBlock(cout), ResidualBlock(cout))


class Dense3D(nn.Module):                         # layers.py:646-726
    def __init__(self, c):
        super().__init__()
        k = {"a": ((3, 3, 1), (1, 1, 0)), "b": ((5, 5, 3), (2, 2, 1)), "c": ((7, 7, 5), (3, 3, 2))}
        self.convs = nn.ModuleDict()
        self.bns = nn.ModuleDict()
        for name, kind in [("1", "a"), ("2", "a"), ("3", "b"), ("4", "c"), ("5", "a"), ("6", "b"), ("7", "c"),
                           ("r1", "a"), ("r2", "b"), ("r3", "c")]:
            self.convs[name] = nn.Conv3d(c, c, k[kind][0], 1, padding=k[kind][1], bias=False)
            self.bns[name] = nn.BatchNorm3d(c)
        self.convs["ch"], self.bns["ch"] = nn.Conv3d(c, c, 1, bias=False), nn.BatchNorm3d(c)

    def forward(self, x):
        f = lambda n, t: F.relu(self.bns[n](self.convs[n](t)))      # noqa: E731
        x1 = f("1", x)
        x2, x3, x4 = f("2", x1), f("3", x1), f("4", x1)
        t1 = x2 + x3 + x4
        s = x1 + x2 + x3 + x4 + f("5", t1) + f("6", t1) + f("7", t1)
        return x1 + f("ch", s) + f("r1", x) + f("r2", x) + f("r3", x)


class DecoderBlock(nn.Module):                    # decoder_v3.py:77-172
    def __init__(self, cin, cout, n_classes):
        super().__init__()
        self.upsample = deconv_block(cin, cout)
        self.resize = nn.Sequential(ME.MinkowskiBatchNorm(cout + 3),
                                    ME.MinkowskiConvolution(cout + 3, cout, kernel_size=1, bias=True, dimension=3))
        self.process = nn.Sequential(ResidualBlock(cout), ResidualBlock(cout), ResidualBlock(cout))
        self.head = ME.MinkowskiConvolution(cout, n_classes, kernel_size=1, bias=True, dimension=3)
        self.pruning = ME.MinkowskiPruning()

    def forward(self, x, shortcut, gmin, gmax):
        d = self.upsample(x)
        c = d.C
        keep = ((c[:, 1] >= gmin[0]) & (c[:, 1] <= gmax[0]) & (c[:, 2] >= gmin[1]) & (c[:, 2] <= gmax[1])
                & (c[:, 3] >= gmin[2]) & (c[:, 3] <= gmax[2]))
        d = self.pruning(d, keep)
        d = ME.SparseTensor(torch.cat([d.F, d.C[:, 1:].float() / d.tensor_stride[0]], 1),
                            coordinate_map_key=d.coordinate_map_key, coordinate_manager=d.coordinate_manager)
        h = self.process(self.resize(d) + shortcut)
        return h, self.head(h)


class OracleNet(nn.Module):
    def __init__(self, f=64, n_classes=20, in_ch=283, caps=(25000, 120000, 400000), queries=100):
        super().__init__()
        self.caps = dict(zip((4, 2, 1), caps))
        self.pp = nn.Sequential(nn.BatchNorm1d(in_ch), nn.Linear(in_ch, 64), nn.BatchNorm1d(64), nn.ReLU(),
                                nn.Linear(64, 128), nn.BatchNorm1d(128), nn.ReLU(), nn.Linear(128, 256),
                                nn.BatchNorm1d(256), nn.ReLU(), nn.Linear(256, f))
        self.stem = ME.MinkowskiConvolution(f, f, kernel_size=1, dimension=3)
        self.s1 = nn.Sequential(ResidualBlock(f), ResidualBlock(f), ResidualBlock(f))
        self.s1s2, self.s2s4, self.s4s8 = down_stage(f, 2 * f), down_stage(2 * f, 4 * f), down_stage(4 * f, 4 * f)
        self.dense = Dense3D(4 * f)
        self.dec = nn.ModuleList([DecoderBlock(4 * f, 4 * f, n_classes), DecoderBlock(4 * f, 2 * f, n_classes),
                                  DecoderBlock(2 * f, f, n_classes)])
        self.voxel_feats = nn.ModuleDict({str(s): nn.Sequential(
            ME.MinkowskiConvolution(c, c, kernel_size=3, dimension=3), ME.MinkowskiBatchNorm(c), ME.MinkowskiReLU(),
            ME.MinkowskiConvolution(c, c, kernel_size=3, bias=True, dimension=3)) for s, c in ((4, 4 * f), (2, 2 * f), (1, f))})
        self.pruning = ME.MinkowskiPruning()
        d = 384
        self.d, self.q = d, queries
        self.query_feat, self.query_embed = nn.Embedding(queries, d), nn.Embedding(queries, d)
        self.input_projs = nn.ModuleList([nn.Linear(4 * f, d), nn.Linear(2 * f, d), nn.Linear(f, d)])
        self.cross = nn.ModuleList([nn.MultiheadAttention(d, 8, batch_first=True) for _ in range(3)])
        self.cross_norm = nn.ModuleList([nn.LayerNorm(d) for _ in range(3)])
        self.selfa = nn.ModuleList([nn.MultiheadAttention(d, 8, batch_first=True) for _ in range(3)])
        self.self_norm = nn.ModuleList([nn.LayerNorm(d) for _ in range(3)])
        self.ffn = nn.ModuleList([nn.Sequential(nn.Linear(d, 1024), nn.ReLU(), nn.Linear(1024, d)) for _ in range(3)])
        self.ffn_norm = nn.ModuleList([nn.LayerNorm(d) for _ in range(3)])
        self.decoder_norm = nn.LayerNorm(d)
        self.class_embed = nn.Linear(d, n_classes + 1)
        self.mask_embed = nn.Sequential(nn.Linear(d, d), nn.ReLU(), nn.Linear(d, d), nn.ReLU(), nn.Linear(d, d))
        self.mask_feat_proj = nn.Linear(f, d)
        self.max_pools = nn.ModuleDict({str(s): ME.MinkowskiMaxPooling(kernel_size=s, stride=s, dimension=3) for s in (2, 4)})

    @staticmethod
    def pe(coords, n):                            # position_encoding.py:90-135
        c = coords.float()
        c = c / (c + 1e-6) * (2 * math.pi)
        t = torch.arange(n, dtype=torch.float32)
        t = 10000.0 ** (2 * torch.div(t, 2, rounding_mode="floor") / n)
        return torch.cat([torch.cat([(c[:, a, None] / t)[:, 0::2].sin(), (c[:, a, None] / t)[:, 1::2].cos()], 1)
                          for a in range(3)], 1)

    def heads(self, out, voxel_feat):
        dn = self.decoder_norm(out)
        return self.class_embed(dn), voxel_feat @ self.mask_embed(dn)[0].t()

    def attn_mask(self, mask_logits, vox, src, scale, lo, hi):     # transformer_predictor_v2.py:220-289 (dense volume)
        keep = ME.SparseTensor((mask_logits.detach().sigmoid() > 0.5).float(), coordinate_map_key=vox.coordinate_map_key,
                               coordinate_manager=vox.coordinate_manager)
        if scale != 1:
            keep = self.max_pools[str(scale)](keep)
        size = [int((h - l) // scale + 1) for l, h in zip(lo, hi)]
        dense = keep.dense(torch.Size([1, keep.F.shape[1], *size]), min_coordinate=torch.IntTensor(list(lo)))[0]
        c = ((src.C[:, 1:] - torch.tensor(lo)) // scale).long()
        m = ~(dense[0, :, c[:, 0], c[:, 1], c[:, 2]].bool())        # [Q,P]
        m[m.all(1)] = False
        return m

    def forward(self, scene, test=False):
        ind = torch.cat([F.pad(c, (1, 0), value=b) for b, c in enumerate(scene["in_coords"])], 0)
        unq, inv = torch.unique(ind, return_inverse=True, dim=0)
        pooled = ME.scatter_max(self.pp(torch.cat(scene["in_feats"], 0)), inv, dim=0)[0]
        x = ME.SparseTensor(pooled, unq.int())
        gmin, gmax = scene["global_min_Cs"], scene["global_max_Cs"]
        lo, hi = [int(v) for v in scene["min_Cs"][0]], [int(v) for v in scene["max_Cs"][0]]
        e1 = self.s1(self.stem(x))
        e2 = self.s1s2(e1)
        e4 = self.s2s4(e2)
        e8 = self.s4s8(e4)
        size = [int(math.ceil((int(h) - int(l) + 1) / 8)) for l, h in zip(gmin, gmax)]
        dense = e8.dense(torch.Size([1, e8.F.shape[1], *size]), min_coordinate=torch.IntTensor([int(v) for v in gmin]))[0]
        sp = ME.to_sparse(self.dense(dense))
        c = sp.C.clone()
        c[:, 1:] = c[:, 1:] * 8 + gmin.int().view(1, -1)
        h = ME.SparseTensor(sp.F, c, tensor_stride=8, coordinate_manager=e8.coordinate_manager)
        sem, xs = {}, {}
        for blk, skip, scale in zip(self.dec, (e4, e2, e1), (4, 2, 1)):
            h, logit = blk(h, skip, gmin, gmax)
            prob, cls = F.softmax(logit.F, -1).max(-1)
            keep = cls != 0
            if not test and int(keep.sum()) > self.caps[scale]:
                pick = torch.multinomial(prob, int(self.caps[scale]), replacement=False)
                keep = torch.zeros_like(keep)
                keep[pick] = True
            h, logit = self.pruning(h, keep), self.pruning(logit, keep)
            sem[scale], xs[scale] = [logit], h
        feats = {}
        for scale, xx in xs.items():
            keep = sem[scale][0].F.max(-1)[1] != 0
            if int(keep.sum()) == 0:
                keep[:1000] = True
            feats[scale] = self.voxel_feats[str(scale)](self.pruning(xx, keep))
        out = self.query_feat.weight.view(1, self.q, self.d)
        qpos = self.query_embed.weight.view(1, self.q, self.d)
        vox = feats[1]
        voxel_feat = self.mask_feat_proj(vox.F) + self.pe(vox.C[:, 1:], self.d // 3)
        cls_l, msk = self.heads(out, voxel_feat)
        classes, masks = [cls_l], [msk]
        for i, s in enumerate((4, 2, 1)):
            src = feats[s]
            kv = (self.input_projs[i](src.F) + self.pe(src.C[:, 1:], self.d // 3)).unsqueeze(0)
            m = self.attn_mask(msk, vox, src, s, lo, hi)
            qn = self.cross_norm[i](out)
            out = qn + self.cross[i](qn + qpos, kv, kv, attn_mask=m.unsqueeze(0).repeat(8, 1, 1))[0]   # blocks.py:82,91
            out = self.self_norm[i](out + self.selfa[i](out + qpos, out + qpos, out)[0])
            out = out + self.ffn[i](self.ffn_norm[i](out))
            cls_l, msk = self.heads(out, voxel_feat)
            classes.append(cls_l)
            masks.append(msk)
        mk = [ME.SparseTensor(t, coordinate_map_key=vox.coordinate_map_key, coordinate_manager=vox.coordinate_manager)
              for t in masks]
        pred = {"query_logits": classes[-1], "voxel_logits": mk[-1],
                "aux_outputs": [{"query_logits": a, "voxel_logits": b} for a, b in zip(classes[:-1], mk[:-1])]}
        return {"sem_logits_at_scales": sem, "panop_predictions": [pred]}


CROP_GRID, CROP_DIV = (64, 64, 32), 16      # THE bounded sample of the benchmark scene: 1/16 of 256x256x32, caps scaled 1/16


def time_crop(repeats: int = 3, warmup: int = 1, occ=0.10, grid=CROP_GRID, div=CROP_DIV):
    """Forward + backward (+ losses) of the CPU restatement on ONE fixed crop of the benchmark scene (the same crop for
    bench.py's cpu_baseline leg and for --impl reference): `warmup` untimed passes, then `repeats` timed passes.
    Returns the median time, the spread and scenes/s scaled to a full scene under the stated assumption that the cost is
    linear in the voxel count (the crop keeps the occupancy, the z extent and the cap-to-voxel ratio of the full scene)."""
    from scene_and_loss import make_scene, total_loss       # the oracle's own copy: nothing of the product on this arm
    torch.manual_seed(0)
    net = OracleNet(caps=tuple(max(8, c // div) for c in (25000, 120000, 400000))).train()
    scene = make_scene(grid, occ, 1, seed=0)
    freq = {f"1_{s}": np.ones(20) for s in (1, 2, 4)}
    times = []
    for i in range(warmup + repeats):
        for q in net.parameters():
            q.grad = None
        t0 = time.time()
        out = net(scene)
        loss = total_loss(out, scene, 20, freq)
        loss.backward()
        if i >= warmup:
            times.append(time.time() - t0)
    times.sort()
    med = times[len(times) // 2]
    return {"scenes_per_s": (1.0 / div) / med, "seconds": round(med, 3), "times": [round(t, 3) for t in times],
            "spread": round((times[-1] - times[0]) / med, 3), "div": div,
            "sample": f"{grid[0]}x{grid[1]}x{grid[2]} crop (1/{div} of a scene) @ {occ:.0%} occ, fwd+bwd+losses, caps scaled "
                      f"1/{div}; {warmup} warm-up + {repeats} timed passes, median; scaled to a full scene assuming cost "
                      f"linear in voxels"}


def time_full_scene(occ=0.10):
    """One full 256x256x32 scene forward + backward (+ losses), no scaling (minutes on 32 cores)."""
    r = time_crop(repeats=1, warmup=0, occ=occ, grid=(256, 256, 32), div=1)
    r["sample"] = f"one full 256x256x32 scene @ {occ:.0%} occ, fwd+bwd+losses, single un-warmed pass"
    return r


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    print(time_full_scene() if "--full" in sys.argv else time_crop())
