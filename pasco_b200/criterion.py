"""The reference's training criterion, term for term ("loss path").

`pasco_b200/losses.py` is the compact loss the benchmark closes its fwd+bwd loop with; THIS module restates what
`Net.step` really optimises and is pinned against the reference's own loss code (tests/test_criterion.py ←
tests/golden/make_golden_loss.py, which runs pasco/loss/* from /root/reference and checks its composition against the
reference's real `Net.step`):

  completion_loss     pasco/loss/losses.py:69-179        class-weighted CE (ignore 255) + Lovász-softmax over ALL rows
                                                         inside the scene bounds (a 255 label is background for every
                                                         class there — `ignores=[255]` never matches a class id)
  lovasz_softmax_flat pasco/loss/lovasz.py:186-219       classes "present", optional ignored class ids; all classes in one
                                                         batched sort instead of the per-class Python loop
  HungarianMatcher    pasco/loss/matcher_sparse.py:69-157 cost = 20·focal + 1·(−p_class) + 1·dice over the voxels that carry a
                                                         target and are not unknown, scaled per target by its class
                                                         weight; SciPy assignment on the host exactly where the
                                                         reference has its `C.cpu()`
  SetCriterion        pasco/loss/criterion_sparse.py:56-411  class CE (per-subnet class weights, unreduced → mean over the
                                                         queries), focal + dice on the matched pairs over the known
                                                         voxels, voxel↔query semantic consistency (CE ignore 0 + Lovász
                                                         ignore class 0 on `semantic_inference_v2` logits) — and EVERY aux
                                                         level is matched again (the `indices` argument of the
                                                         reference is never read)
  training_loss       pasco/models/net_panoptic_sparse.py:355-447  the total of `Net.step`: (CE + Lovász)·occ_weight +
                                                         (dice + CE + mask)·panop_weight/M + Σ aux terms/M; the
                                                         main level's ssc terms are computed by the reference but never
                                                         added (the two accumulators stay 0.0) — skipped here

Device-agnostic torch code: a sparse tensor is anything with `.F` [N, C] and `.C` [N, 4] (b, x, y, z) — a
`pasco_b200.me.SparseTensor` on the GPU, a `Rows` tuple in the CPU tests."""
from __future__ import annotations

from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

WEIGHTS = {"ssc_ce": 0.3, "ssc_lovasz": 1.0, "loss_ce": 2.0, "loss_mask": 20.0, "loss_dice": 1.0}   # net_panoptic_sparse.py:139-151


class Rows(NamedTuple):
    F: torch.Tensor
    C: torch.Tensor


def lovasz_softmax_flat(logits: torch.Tensor, labels: torch.Tensor, ignores: Sequence[int] = ()) -> torch.Tensor:
    """lovasz.py:186-219 with classes="present": mean over the classes that occur in `labels` and are not in `ignores` of
    the Lovász extension of the Jaccard loss.  One batched sort over [C, P] (no per-class loop, no host round trip)."""
    if logits.shape[0] == 0:
        return logits.sum() * 0.0
    probs = F.softmax(logits, dim=1)
    n_cls = probs.shape[1]
    cls = torch.arange(n_cls, device=labels.device).view(-1, 1)
    fg = (labels.view(1, -1) == cls).to(probs.dtype)                 # [C, P]
    keep = fg.sum(1) > 0
    for c in ignores:
        if 0 <= c < n_cls:
            keep[c] = False
    err = (fg - probs.t()).abs()
    err_sorted, perm = torch.sort(err, dim=1, descending=True)
    fg_sorted = fg.gather(1, perm)
    gts = fg_sorted.sum(1, keepdim=True)
    inter = gts - fg_sorted.cumsum(1)
    union = gts + (1.0 - fg_sorted).cumsum(1)
    jac = 1.0 - inter / union                                        # lovasz_grad, lovasz.py:19-31
    jac = torch.cat([jac[:, :1], jac[:, 1:] - jac[:, :-1]], 1)
    per_class = (err_sorted * jac).sum(1)
    w = keep.to(probs.dtype)
    return (per_class * w).sum() / w.sum().clamp(min=1.0)


def label_weights(class_frequencies, power: float = 1.0 / 3.0) -> torch.Tensor:
    """losses.py:134-140 (SemanticKITTI, power 1/3) / :80-86 (KITTI-360, power 1/1.5)."""
    fr = np.asarray(class_frequencies, dtype=np.float64)
    w = fr / fr.sum()
    return torch.from_numpy(np.power(np.amax(w) / w, power))


def completion_loss(sem_labels: Dict[str, torch.Tensor], sem_logits_at_scales: Dict[int, list], min_Cs, max_Cs,
                    class_frequencies: Dict[str, np.ndarray], power: float = 1.0 / 3.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """losses.py:121-179 → (CE, Lovász), each the mean over every (scale, subnet) pair with at least one row inside the
    subnet's bounds (misc.py:16-27)."""
    ces, lovs = [], []
    for scale, per_subnet in sem_logits_at_scales.items():
        targets = sem_labels[f"1_{scale}"]
        for m in range(len(targets)):
            lg = per_subnet[m]
            dev = lg.F.device
            w = label_weights(class_frequencies[f"1_{scale}"], power).to(device=dev, dtype=lg.F.dtype)
            lo, hi = torch.as_tensor(min_Cs[m]).to(dev).view(1, 3), torch.as_tensor(max_Cs[m]).to(dev).view(1, 3)
            xyz = lg.C[:, 1:]
            inside = ((xyz >= lo) & (xyz <= hi)).all(1)
            if not bool(inside.any()):
                continue
            cell = torch.div(xyz[inside].long() - lo.long(), scale, rounding_mode="floor")
            t = targets[m].to(dev)[cell[:, 0], cell[:, 1], cell[:, 2]].long()
            logits = lg.F[inside]
            ces.append(F.cross_entropy(logits, t, weight=w, ignore_index=255))
            lovs.append(lovasz_softmax_flat(logits, t, ignores=(255,)))
    if not ces:
        return torch.zeros(()), torch.zeros(())
    return torch.stack(ces).mean(), torch.stack(lovs).mean()


def sigmoid_focal(logits: torch.Tensor, targets: torch.Tensor, alpha: float = 0.25, gamma: float = 2.0) -> torch.Tensor:
    """losses.py:45-66, elementwise."""
    p = logits.sigmoid()
    ce = F.binary_cross_entropy_with_logits(logits, targets, reduction="none")
    p_t = p * targets + (1 - p) * (1 - targets)
    return (alpha * targets + (1 - alpha) * (1 - targets)) * ce * (1 - p_t) ** gamma


def dice(logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """losses.py:27-42: per column (sums over the voxel dimension 0)."""
    s = logits.sigmoid()
    return 1 - (2 * (s * targets).sum(0) + 1) / (s.sum(0) + targets.sum(0) + 1)


class HungarianMatcher:
    """matcher_sparse.py:69-157.  __call__(query_logits [Q, K+1], mask_logits [N, Q], labels [T], masks [N, T] (0/1),
    class_weight [K+1], unknown [N] bool) → (query idx, target idx), int64 on the host."""

    def __init__(self, cost_class: float = 1.0, cost_mask: float = 20.0, cost_dice: float = 1.0):
        self.cost_class, self.cost_mask, self.cost_dice = cost_class, cost_mask, cost_dice

    @torch.no_grad()
    def cost(self, query_logits, mask_logits, labels, masks, class_weight, unknown) -> torch.Tensor:
        prob = query_logits.softmax(-1)
        ids = labels.long()
        out = mask_logits.t()                                         # [Q, N]
        tgt = masks.t().to(out)                                       # [T, N]
        valid = (tgt.sum(0) > 0) & ~unknown
        out, tgt = out[:, valid], tgt[:, valid]
        cost_class = -prob[:, ids]
        s = out.sigmoid()
        cost_dice = 1 - (2 * s @ tgt.t() + 1) / (s.sum(-1)[:, None] + tgt.sum(-1)[None, :] + 1)
        if out.shape[1] != 0:
            bce_pos = F.binary_cross_entropy_with_logits(out, torch.ones_like(out), reduction="none")
            bce_neg = F.binary_cross_entropy_with_logits(out, torch.zeros_like(out), reduction="none")
            pos = 0.25 * (1 - s) ** 2 * bce_pos
            neg = 0.75 * s ** 2 * bce_neg
            cost_mask = (pos @ tgt.t() + neg @ (1 - tgt).t()) / out.shape[1]
        else:
            cost_mask = torch.zeros_like(cost_dice)
        c = self.cost_mask * cost_mask + self.cost_class * cost_class + self.cost_dice * cost_dice
        return c * class_weight.to(c)[ids][None, :]

    def __call__(self, query_logits, mask_logits, labels, masks, class_weight, unknown):
        c = self.cost(query_logits, mask_logits, labels, masks, class_weight, unknown)
        i, j = linear_sum_assignment(c.float().cpu().numpy())
        return torch.as_tensor(i, dtype=torch.int64), torch.as_tensor(j, dtype=torch.int64)


def semantic_inference(mask_probs: torch.Tensor, query_logits: torch.Tensor) -> Optional[torch.Tensor]:
    """helper.py:7-39 (`semantic_inference_v2`): voxel class logits as the mask-probability-weighted mean of the class logits
    of the queries whose arg-max is not the dustbin; None when every query is dustbin."""
    n_classes = query_logits.shape[-1] - 1
    keep = query_logits.softmax(-1).argmax(-1) != n_classes
    if not bool(keep.any()):
        return None
    p = mask_probs[:, keep] + 1e-8
    return (p / p.sum(1, keepdim=True)) @ query_logits[keep, :-1]


class SetCriterion:
    """criterion_sparse.py:19-411 for one scene per call (the reference's `bs` is 1 per subnet, net_panoptic_sparse.py:417).

    class_weights: one [K+1] tensor per subnet (scripts/train.py:117-123); compl_labelweights [K] (train.py:125-128)."""

    def __init__(self, num_classes: int, class_weights: Sequence[torch.Tensor], compl_labelweights: torch.Tensor,
                 weights: Optional[Dict[str, float]] = None, matcher: Optional[HungarianMatcher] = None):
        self.num_classes = num_classes
        self.class_weights = list(class_weights)
        self.compl_labelweights = compl_labelweights
        self.w = dict(WEIGHTS if weights is None else weights)
        self.matcher = matcher or HungarianMatcher(1.0, self.w["loss_mask"], self.w["loss_dice"])
        self.last_indices: List[Tuple[torch.Tensor, torch.Tensor]] = []

    def level_losses(self, voxel_logits, query_logits, labels, masks_dense, semantic_label, unknown_dense, i_infer: int,
                     min_C, with_ssc: bool = True) -> Dict[str, torch.Tensor]:
        """compute_losses (criterion_sparse.py:239-351) for one output level.  voxel_logits: rows [N, Q] with absolute
        coordinates; query_logits [1, Q, K+1]; masks_dense bool [T, X, Y, Z], semantic_label [X, Y, Z] and unknown_dense
        bool [X, Y, Z] indexed by coordinate − min_C."""
        Fv, dev = voxel_logits.F, voxel_logits.F.device
        cw = self.class_weights[i_infer].to(dev)
        q = query_logits[0]
        cell = voxel_logits.C[:, 1:].long() - torch.as_tensor(min_C).to(dev).long().view(1, 3)
        x, y, z = cell[:, 0], cell[:, 1], cell[:, 2]
        masks = masks_dense.to(dev)[:, x, y, z].t().to(Fv.dtype)       # [N, T]
        unknown = unknown_dense.to(dev)[x, y, z]
        labels = labels.to(dev)
        qi, tj = self.matcher(q, Fv.detach(), labels, masks, cw, unknown)
        self.last_indices.append((qi, tj))
        qi, tj = qi.to(dev), tj.to(dev)
        # loss_labels (:56-81): unreduced weighted CE over the Q queries, then a plain mean
        target_cls = torch.full((q.shape[0],), self.num_classes, dtype=torch.int64, device=dev)
        target_cls[qi] = labels[tj].long()
        loss_ce = (F.cross_entropy(q, target_cls, cw.to(q.dtype), reduction="none") * self.w["loss_ce"]).mean()
        # loss_masks (:83-112): matched pairs over the known voxels, weighted by the class weight of each target
        tw = cw[labels[tj].long()]
        known = ~unknown
        src, tgt = Fv[:, qi][known], masks[:, tj][known]
        loss_mask = ((sigmoid_focal(src, tgt) * tw.unsqueeze(0)).mean(0) * self.w["loss_mask"]).mean()
        loss_dice = (dice(src, tgt) * tw * self.w["loss_dice"]).mean()
        out = {"loss_ce": loss_ce, "loss_mask": loss_mask, "loss_dice": loss_dice}
        if with_ssc:
            out.update(self.ssc_losses(Fv, q, cell, semantic_label))
        return out

    def ssc_losses(self, mask_logits, q, cell, semantic_label) -> Dict[str, torch.Tensor]:
        """compute_ssc_sparse_loss (:182-210), weighted (:337-342)."""
        dev = mask_logits.device
        zero = mask_logits.sum() * 0.0
        logits = semantic_inference(mask_logits.sigmoid(), q)
        if logits is None:
            return {"ssc_ce_loss": zero, "ssc_lovasz_loss": zero}
        t = semantic_label.to(dev)[cell[:, 0], cell[:, 1], cell[:, 2]].long()
        ok = t != 255
        logits, t = logits[ok], t[ok]
        ce = F.cross_entropy(logits, t, weight=self.compl_labelweights.to(logits), ignore_index=0)
        lov = lovasz_softmax_flat(logits, t, ignores=(0,))
        return {"ssc_ce_loss": ce * self.w["ssc_ce"], "ssc_lovasz_loss": lov * self.w["ssc_lovasz"]}

    def __call__(self, pred: Dict, labels, masks_dense, semantic_label, unknown_dense, i_infer: int, min_C,
                 main_ssc: bool = False) -> Dict[str, torch.Tensor]:
        """forward (:353-411): main level + one re-matched pass per aux level (keys `<term>_level<i>`).  main_ssc=False
        skips the main level's ssc terms, which `Net.step` computes and then drops."""
        self.last_indices = []
        losses = self.level_losses(pred["voxel_logits"], pred["query_logits"], labels, masks_dense, semantic_label,
                                   unknown_dense, i_infer, min_C, with_ssc=main_ssc)
        aux = {}
        for i, a in enumerate(pred.get("aux_outputs", [])):
            for k, v in self.level_losses(a["voxel_logits"], a["query_logits"], labels, masks_dense, semantic_label,
                                          unknown_dense, i_infer, min_C).items():
                aux[f"{k}_level{i}"] = v
        losses["loss_aux"] = aux
        return losses


def training_loss(out: Dict, batch: Dict, criterion: SetCriterion, class_frequencies, occ_weight: float = 1.0,
                  panop_weight: float = 1.0, power: float = 1.0 / 3.0):
    """The loss of `Net.step` (net_panoptic_sparse.py:355-447) from a forward output dict (`sem_logits_at_scales`,
    `panop_predictions`) and the reference's batch fields: sem_labels {"1_s": [M, X/s, Y/s, Z/s]}, min_Cs / max_Cs,
    semantic_label [M, X, Y, Z], mask_label [{"labels" [T], "masks" [T, X, Y, Z]}], geo_labels["1_1"] [M, X, Y, Z]
    (255 = unknown).  Returns (total, terms)."""
    ce, lov = completion_loss(batch["sem_labels"], out["sem_logits_at_scales"], batch["min_Cs"], batch["max_Cs"],
                              class_frequencies, power)
    total = (ce + lov) * occ_weight
    terms = {"compl_ce_loss": ce, "compl_lovasz_loss": lov}
    preds = out.get("panop_predictions")
    if not preds:
        return total, terms
    n = len(preds)
    for m, pred in enumerate(preds):
        ml = batch["mask_label"][m]
        unknown = batch["geo_labels"]["1_1"][m] == 255
        ls = criterion(pred, ml["labels"], ml["masks"], batch["semantic_label"][m], unknown, m, batch["min_Cs"][m])
        for k in ("loss_ce", "loss_mask", "loss_dice"):
            terms[k] = terms.get(k, 0.0) + ls[k] / n
        for k, v in ls["loss_aux"].items():
            terms[k] = terms.get(k, 0.0) + v / n
    total = total + (terms["loss_dice"] + terms["loss_ce"] + terms["loss_mask"]) * panop_weight
    for k, v in terms.items():
        if "_level" in k:
            total = total + v
    return total, terms
