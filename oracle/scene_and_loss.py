"""Synthetic scene + training losses for the CPU arm (`oracle/net_oracle.py`: bench.py's `cpu_baseline` leg and
`--impl reference`).  TEST INFRASTRUCTURE — deliberately a SEPARATE COPY of pasco_b200/synthetic.py and pasco_b200/losses.py
so that the reference arm imports nothing from the product package.  tests/test_losses.py checks
that the two copies agree on the same inputs.

  make_scene          SURVEY.md §8d synthetic scene (Bernoulli occupancy, one point per voxel, N(0,1) features)
  completion_loss     pasco/loss/losses.py:124-179      class-weighted CE + Lovász-softmax at scales 1, 2, 4
  panoptic_set_loss   pasco/loss/criterion_sparse.py:19-411 + matcher_sparse.py:69-157   Hungarian set loss
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment


def clustered_occupancy(grid: Sequence[int], occ: float, g: torch.Generator) -> torch.Tensor:
    """Lidar-like occupancy with the same voxel budget as the Bernoulli scene: a noisy ground slab plus random boxes
    (walls, cars, vegetation blobs), thinned or topped up to exactly round(occ·X·Y·Z) voxels (the second
    measurement point — real scenes are clustered, so tiles see more neighbours and fewer rows have none)."""
    X, Y, Z = grid
    target = int(round(occ * X * Y * Z))
    dens = torch.zeros(X, Y, Z)
    dens[:, :, max(0, Z // 8 - 1): Z // 8 + 2] = 0.6                                   # ground slab
    n_boxes = max(4, (X * Y) // 600)
    for _ in range(n_boxes):
        sx, sy = int(torch.randint(6, max(7, X // 5), (1,), generator=g)), int(torch.randint(6, max(7, Y // 5), (1,), generator=g))
        sz = int(torch.randint(2, max(3, Z // 2), (1,), generator=g))
        x0, y0 = int(torch.randint(0, X - sx + 1, (1,), generator=g)), int(torch.randint(0, Y - sy + 1, (1,), generator=g))
        z0 = int(torch.randint(0, max(1, Z - sz + 1), (1,), generator=g))
        dens[x0:x0 + sx, y0:y0 + sy, z0:z0 + sz] = torch.maximum(dens[x0:x0 + sx, y0:y0 + sy, z0:z0 + sz], torch.tensor(0.5))
    score = dens + 0.45 * torch.rand(X, Y, Z, generator=g)          # structure first, uniform noise breaks ties / tops up
    thr = torch.topk(score.view(-1), target).values[-1]
    return score >= thr


def make_scene(grid: Sequence[int] = (256, 256, 32), occ: float = 0.10, n_infers: int = 1, in_ch: int = 283,
               n_classes: int = 20, seed: int = 0, n_masks: int = 10, clustered: bool = False) -> Dict:
    g = torch.Generator().manual_seed(seed)
    X, Y, Z = grid
    b: Dict = {"in_feats": [], "in_coords": [], "min_Cs": [], "max_Cs": [], "Ts": []}
    for _ in range(n_infers):
        o = clustered_occupancy(grid, occ, g) if clustered else torch.rand(X, Y, Z, generator=g) < occ
        c = torch.nonzero(o).int()
        b["in_coords"].append(c)
        b["in_feats"].append(torch.randn(c.shape[0], in_ch, generator=g))
        b["min_Cs"].append(torch.tensor([0, 0, 0]))
        b["max_Cs"].append(torch.tensor([X - 1, Y - 1, Z - 1]))
        b["Ts"].append(torch.eye(4))
    b["global_min_Cs"] = torch.tensor([0, 0, 0])
    b["global_max_Cs"] = torch.tensor([X - 1, Y - 1, Z - 1])
    sem = {}
    for s in (1, 2, 4):
        lab = torch.randint(1, n_classes, (n_infers, X // s, Y // s, Z // s), generator=g)
        lab[torch.rand(lab.shape, generator=g) < 0.9] = 0
        sem[f"1_{s}"] = lab.to(torch.uint8)
    b["sem_labels"] = sem
    # instance masks: random boxes (dense bool [n_masks, X, Y, Z]) with a thing/stuff class each
    boxes = []
    for _ in range(n_masks):
        lo = [int(torch.randint(0, d - d // 4, (1,), generator=g)) for d in (X, Y, Z)]
        boxes.append((lo, [min(d, l + max(2, d // 4)) for l, d in zip(lo, (X, Y, Z))]))
    b["mask_boxes"] = boxes
    b["mask_classes"] = torch.randint(1, n_classes, (n_masks,), generator=g)
    return b


def lovasz_softmax_present(probs: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """Lovász-softmax over the classes present in `labels` (Berman et al. 2018).  All classes are sorted in one
    batched sort along dim 0 (no per-class loop, no host synchronisation); absent classes get weight 0."""
    n_cls = probs.shape[1]
    if probs.shape[0] == 0:
        return probs.sum() * 0
    cls = torch.arange(n_cls, device=labels.device).view(-1, 1)
    fg = (labels.view(1, -1) == cls).to(probs.dtype)                # [C,P]: class-major so that sort / cumsum run along
    present = (fg.sum(1) > 0).to(probs.dtype)                       # the contiguous dimension
    err = (fg - probs.t()).abs()
    err_sorted, perm = torch.sort(err, dim=1, descending=True)
    fg_sorted = fg.gather(1, perm)
    gts = fg_sorted.sum(1, keepdim=True)
    inter = gts - fg_sorted.cumsum(1)
    union = gts + (1 - fg_sorted).cumsum(1)
    jac = 1.0 - inter / union
    jac = torch.cat([jac[:, :1], jac[:, 1:] - jac[:, :-1]], 1)
    per_class = (err_sorted * jac).sum(1)
    return (per_class * present).sum() / present.sum().clamp(min=1)


def completion_loss(sem_logits_at_scales: Dict[int, list], sem_labels: Dict[str, torch.Tensor], min_Cs,
                    class_frequencies) -> torch.Tensor:
    ces, lovs = [], []
    for scale, per_subnet in sem_logits_at_scales.items():
        fr = np.asarray(class_frequencies[f"1_{scale}"], dtype=np.float64)
        w = fr / fr.sum()
        w = torch.from_numpy(np.power(np.amax(w) / w, 1 / 3.0)).float()
        for m, lg in enumerate(per_subnet):
            if lg.F.shape[0] == 0:
                continue
            c = (lg.C[:, 1:].long() - min_Cs[m].to(lg.C.device).view(1, 3)) // scale
            tgt = sem_labels[f"1_{scale}"][m]
            inside = ((c >= 0).all(1) & (c[:, 0] < tgt.shape[0]) & (c[:, 1] < tgt.shape[1]) & (c[:, 2] < tgt.shape[2]))
            c, logits = c[inside], lg.F[inside]
            t = tgt[c[:, 0], c[:, 1], c[:, 2]].long()
            valid = t != 255
            ces.append(F.cross_entropy(logits, t, weight=w.to(logits), ignore_index=255))
            lovs.append(lovasz_softmax_present(F.softmax(logits[valid], 1), t[valid]))
    if not ces:
        return torch.zeros((), requires_grad=True)
    return torch.stack(ces).mean() + torch.stack(lovs).mean()


def _focal(logits, targets, alpha=0.25, gamma=2.0):
    p = torch.sigmoid(logits)
    ce = F.binary_cross_entropy_with_logits(logits, targets, reduction="none")
    pt = p * targets + (1 - p) * (1 - targets)
    return (alpha * targets + (1 - alpha) * (1 - targets)) * ce * (1 - pt) ** gamma


@torch.no_grad()
def hungarian(query_logits, mask_logits, tgt_cls, tgt_masks, w_class=1.0, w_mask=20.0, w_dice=1.0):
    """query_logits [Q,K+1], mask_logits [P,Q], tgt_masks [T,P] → (query idx, target idx).  CPU crossing
    exactly where the reference has one (matcher_sparse.py:151)."""
    prob = query_logits.softmax(-1)
    out = mask_logits.t().float()                                   # [Q,P]
    # the three [Q,P]x[P,T] cost GEMMs only rank assignments: TF32 is allowed here even in the fp32 parity mode (the fp32
    # CUDA-core library GEMMs took 3 ms per step for 0.8 GFLOP because of the P = 400 k reduction dimension)
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        return _hungarian_costs(prob, out, tgt_cls, tgt_masks.float(), w_class, w_mask, w_dice)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32


def _hungarian_costs(prob, out, tgt_cls, tgt_masks, w_class, w_mask, w_dice):
    cost_class = -prob[:, tgt_cls]
    sig = out.sigmoid()
    num = 2 * sig @ tgt_masks.t()
    den = sig.sum(-1)[:, None] + tgt_masks.sum(-1)[None, :]
    cost_dice = 1 - (num + 1) / (den + 1)
    P = max(out.shape[1], 1)
    pos = _focal(out, torch.ones_like(out))
    neg = _focal(out, torch.zeros_like(out))
    cost_mask = (pos @ tgt_masks.t() + neg @ (1 - tgt_masks).t()) / P
    C = w_mask * cost_mask + w_class * cost_class + w_dice * cost_dice
    i, j = linear_sum_assignment(C.float().cpu().numpy())
    return torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)


def panoptic_set_loss(pred: Dict, tgt_cls: torch.Tensor, tgt_masks: torch.Tensor, n_classes: int, eos_coef=0.1,
                      w_ce=2.0, w_mask=20.0, w_dice=1.0) -> torch.Tensor:
    """pred = one entry of panop_predictions; tgt_masks float [T,P] sampled at pred['voxel_logits'].C."""
    levels = [(pred["query_logits"], pred["voxel_logits"].F)] + \
             [(a["query_logits"], a["voxel_logits"].F) for a in pred.get("aux_outputs", [])]
    dev = levels[0][0].device
    qi, tj = hungarian(levels[0][0][0], levels[0][1], tgt_cls, tgt_masks)
    qi, tj = qi.to(dev), tj.to(dev)
    # the class head always has 20 + 1 outputs (the reference builds TransformerPredictor without num_classes), the last one
    # being "no object" — also for KITTI-360's 19 classes
    n_logits = levels[0][0].shape[-1]
    empty_w = torch.ones(n_logits, device=dev)
    empty_w[-1] = eos_coef
    total = 0.0
    for qlog, mlog in levels:
        q = qlog[0]
        target_cls = torch.full((q.shape[0],), n_logits - 1, dtype=torch.int64, device=dev)
        target_cls[qi] = tgt_cls[tj]
        l_ce = F.cross_entropy(q, target_cls, empty_w)
        src = mlog.t()[qi]                                          # [T',P]
        t = tgt_masks[tj]
        T = max(len(qi), 1)
        l_mask = _focal(src, t).mean(1).sum() / T
        sig = src.sigmoid()
        l_dice = (1 - (2 * (sig * t).sum(1) + 1) / (sig.sum(1) + t.sum(1) + 1)).sum() / T
        total = total + w_ce * l_ce + w_mask * l_mask + w_dice * l_dice
    return total


def masks_at(coords: torch.Tensor, boxes) -> torch.Tensor:
    """Sample box masks (pasco_b200.synthetic.make_scene) at voxel coordinates → float [T,P]."""
    c = coords[:, 1:].unsqueeze(0)                                              # [1,P,3]
    lo = torch.as_tensor([b[0] for b in boxes], device=coords.device, dtype=coords.dtype).view(-1, 1, 3)
    hi = torch.as_tensor([b[1] for b in boxes], device=coords.device, dtype=coords.dtype).view(-1, 1, 3)
    return ((c >= lo) & (c < hi)).all(-1).float()


def total_loss(out: Dict, scene: Dict, n_classes: int, class_frequencies) -> torch.Tensor:
    loss = completion_loss(out["sem_logits_at_scales"], scene["sem_labels"], scene["min_Cs"], class_frequencies)
    for m, pred in enumerate(out.get("panop_predictions", [])):
        tm = masks_at(pred["voxel_logits"].C, scene["mask_boxes"])
        loss = loss + panoptic_set_loss(pred, scene["mask_classes"].to(tm.device), tm, n_classes) / len(out["panop_predictions"])
    return loss
