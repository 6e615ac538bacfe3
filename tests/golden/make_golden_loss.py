"""Golden vectors for pasco_b200/criterion.py — BUILD CONTAINER ONLY: runs the reference's OWN loss code from
/root/reference on the CPU (MinkowskiEngine = the oracle shim):

  * pasco/loss/criterion_sparse.py:19-411   SetCriterion.forward (main level + 3 aux levels, every term)
  * pasco/loss/matcher_sparse.py:69-157     HungarianMatcher (assignment indices)
  * pasco/loss/losses.py:69-179             compute_sem_compl_loss / compute_sem_compl_loss_kitti360
  * pasco/loss/lovasz.py:186-219            lovasz_softmax_flat (through the two above)

on a synthetic M = 2 prediction set (non-zero min_C, unknown voxels, 255 labels, rows outside the scene bounds, dustbin
queries) and composes the total exactly like Net.step (net_panoptic_sparse.py:355-447).  `--check-step` additionally runs
the reference's real `Net.step(batch, "train")` at 32x32x8 on the oracle and asserts that this script's composition of
the same criterion outputs equals the loss the reference returns (nothing is stored from that run).

    python tests/golden/make_golden_loss.py [--check-step]
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), os.path.join(ROOT, "compat"), "/root/reference", ROOT,
                os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "tools"), HERE]

import numpy as np  # noqa: E402
import torch  # noqa: E402

GRID, MIN_C, Q, K, M, T, LEVELS = (32, 32, 8), (8, -4, 0), 24, 20, 2, 9, 4
WEIGHTS = {"ssc_ce": 0.3, "ssc_lovasz": 1.0, "loss_ce": 2.0, "loss_mask": 20.0, "loss_dice": 1.0}


def synthetic(seed=0):
    """Everything the criterion consumes, as plain tensors (the test rebuilds the same from the .npz)."""
    g = torch.Generator().manual_seed(seed)
    X, Y, Z = GRID
    minc = torch.tensor(MIN_C)
    d = {"min_C": minc, "max_C": minc + torch.tensor(GRID) - 1}
    # labels: blobs of classes 1..19, a slab of unknown (255), rest empty (0)
    sem = torch.zeros(M, X, Y, Z, dtype=torch.uint8)
    masks, labels = [], []
    for m in range(M):
        mk, lb = [], []
        for t in range(T):
            lo = [int(torch.randint(0, s - 3, (1,), generator=g)) for s in GRID]
            hi = [min(s, l + int(torch.randint(3, max(4, s // 2), (1,), generator=g))) for l, s in zip(lo, GRID)]
            box = torch.zeros(X, Y, Z, dtype=torch.bool)
            box[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = True
            c = 1 + (3 * t + m) % (K - 1)
            sem[m][box] = c
            mk.append(box)
            lb.append(c)
        sem[m, :, Y - 3:, :] = 255
        masks.append(torch.stack(mk))
        labels.append(torch.tensor(lb, dtype=torch.uint8))
    d["semantic_label"], d["masks"], d["labels"] = sem, torch.stack(masks), torch.stack(labels)
    d["class_weights"] = torch.ones(M, K + 1)
    d["class_weights"][:, 0] = 0.1
    d["class_weights"][:, -1] = 0.1
    d["class_weights"][1, 1:K] = 0.5 + torch.rand(K - 1, generator=g)          # subnet 1: non-uniform (exercises i_infer)
    freq = {f"1_{s}": (1.0 + 50.0 * torch.rand(K, generator=g)).double().numpy() for s in (1, 2, 4)}
    d["class_frequencies"] = np.stack([freq[f"1_{s}"] for s in (1, 2, 4)])
    w = freq["1_1"] / freq["1_1"].sum()
    d["compl_labelweights"] = torch.from_numpy(np.power(np.amax(w) / w, 1 / 3.0)).float()
    # predictions
    for m in range(M):
        occ = torch.rand(X, Y, Z, generator=g) < 0.30
        c = torch.nonzero(occ).int() + minc.int().view(1, 3)
        c = torch.cat([torch.zeros(c.shape[0], 1, dtype=torch.int32), c], 1)
        d[f"vox_C_m{m}"] = c
        inside = d["masks"][m][:, (c[:, 1] - MIN_C[0]).long(), (c[:, 2] - MIN_C[1]).long(), (c[:, 3] - MIN_C[2]).long()].t().float()
        for lv in range(LEVELS):
            perm = torch.randperm(Q, generator=g)
            f = 1.5 * torch.randn(c.shape[0], Q, generator=g) - 1.0
            f[:, perm[:T]] += (3.0 - 0.5 * lv) * inside                      # some queries roughly follow a target mask
            ql = torch.randn(1, Q, K + 1, generator=g)
            ql[0, perm[:T], d["labels"][m].long()] += 2.0
            ql[0, perm[T + 4:], K] += 4.0                                   # most of the rest: dustbin
            d[f"vox_F_m{m}_l{lv}"], d[f"query_m{m}_l{lv}"] = f, ql
        for s in (1, 2, 4):
            # semantic logits at stride s: absolute coordinates are multiples of s; a margin of rows lies outside the scene
            lo = (minc // s) * s - s
            n = [(GRID[i] + 2 * s) // s + 1 for i in range(3)]
            cells = torch.nonzero(torch.rand(*n, generator=g) < (0.5 if s > 1 else 0.25)).int() * s + lo.int().view(1, 3)
            cs = torch.cat([torch.zeros(cells.shape[0], 1, dtype=torch.int32), cells], 1)
            d[f"sem{s}_C_m{m}"], d[f"sem{s}_F_m{m}"] = cs, torch.randn(cs.shape[0], K, generator=g)
    for s in (1, 2, 4):
        lab = torch.randint(0, K, (M, X // s, Y // s, Z // s), generator=g)
        lab[torch.rand(lab.shape, generator=g) < 0.5] = 0
        lab[torch.rand(lab.shape, generator=g) < 0.1] = 255
        d[f"sem_labels_1_{s}"] = lab.to(torch.uint8)
    return d


def reference_losses(d):
    """The reference's own functions on the synthetic set; total composed as Net.step does."""
    import MinkowskiEngine as ME
    from pasco.loss.criterion_sparse import SetCriterion
    from pasco.loss.matcher_sparse import HungarianMatcher
    from pasco.loss.losses import compute_sem_compl_loss, compute_sem_compl_loss_kitti360
    matcher = HungarianMatcher(cost_class=1.0, cost_mask=WEIGHTS["loss_mask"], cost_dice=WEIGHTS["loss_dice"])
    crit = SetCriterion(K, matcher=matcher, weight_dict=WEIGHTS, eos_coef=0.1, class_weights=list(d["class_weights"]),
                        compl_labelweights=d["compl_labelweights"])
    leaves, res = {}, {}

    def leaf(name):
        leaves[name] = d[name].clone().requires_grad_(True)
        return leaves[name]

    sem_at = {s: [ME.SparseTensor(leaf(f"sem{s}_F_m{m}"), d[f"sem{s}_C_m{m}"]) for m in range(M)] for s in (1, 2, 4)}
    sem_labels = {f"1_{s}": d[f"sem_labels_1_{s}"] for s in (1, 2, 4)}
    freqs = {f"1_{s}": d["class_frequencies"][i] for i, s in enumerate((1, 2, 4))}
    min_Cs, max_Cs = [d["min_C"]] * M, [d["max_C"]] * M
    ce, lov = compute_sem_compl_loss(sem_labels, sem_at, min_Cs, max_Cs, freqs)
    ce360, lov360 = compute_sem_compl_loss_kitti360(sem_labels, sem_at, min_Cs, max_Cs, freqs)
    res["compl_ce"], res["compl_lovasz"] = ce.item(), lov.item()
    res["compl_ce_kitti360"], res["compl_lovasz_kitti360"] = ce360.item(), lov360.item()
    total = (ce + lov) * 1.0
    loss_ce = loss_mask = loss_dice = 0.0
    aux = {}
    for m in range(M):
        def level(lv):
            return {"voxel_logits": ME.SparseTensor(leaf(f"vox_F_m{m}_l{lv}"), d[f"vox_C_m{m}"]),
                    "query_logits": leaf(f"query_m{m}_l{lv}")}
        pred = level(0)
        pred["aux_outputs"] = [level(lv) for lv in range(1, LEVELS)]
        target = [{"labels": d["labels"][m], "masks": d["masks"][m]}]
        unknown = (d["semantic_label"][m] == 255).unsqueeze(0)
        # record the assignments (the criterion calls the matcher again with the same inputs → same result)
        for lv in range(LEVELS):
            lvl = pred if lv == 0 else pred["aux_outputs"][lv - 1]
            c = lvl["voxel_logits"].C.clone()
            c[:, 1:] -= d["min_C"].int().view(1, 3)
            tm = d["masks"][m][:, c[:, 1].long(), c[:, 2].long(), c[:, 3].long()].t()
            i, j = matcher({"query_logits": lvl["query_logits"][0], "voxel_logits": ME.SparseTensor(lvl["voxel_logits"].F.detach(), c)},
                           {"labels": d["labels"][m], "masks": ME.SparseTensor(tm.float(), c)}, d["class_weights"][m], unknown)
            res[f"match_m{m}_l{lv}"] = torch.stack([i, j]).numpy()
        out = crit(None, pred, target, d["semantic_label"][m].unsqueeze(0), unknown, m, [], d["min_C"])
        for k in ("loss_ce", "loss_mask", "loss_dice", "ssc_ce_loss", "ssc_lovasz_loss"):
            res[f"{k}_m{m}"] = float(out[k])
        for k, v in out["loss_aux"].items():
            res[f"{k}_m{m}"] = float(v)
            aux[k] = aux.get(k, 0.0) + v / M
        loss_ce = loss_ce + out["loss_ce"] / M
        loss_mask = loss_mask + out["loss_mask"] / M
        loss_dice = loss_dice + out["loss_dice"] / M
    total = total + (loss_dice + loss_ce + loss_mask) * 1.0 + (0.0 + 0.0)       # net_panoptic_sparse.py:472-474
    for k in aux:
        total = total + aux[k]
    res["total"] = total.item()
    total.backward()
    grads = {f"grad_{n}": (t.grad if t.grad is not None else torch.zeros_like(t)).numpy() for n, t in leaves.items()}
    return res, grads


def check_against_net_step():
    """The reference's real training step on the oracle (32x32x8, M=2) vs this script's composition of its outputs."""
    import MinkowskiEngine as ME
    from recipe import fill_state_dict
    from run_reference_on_oracle import build_net, synthetic_batch
    from pasco.loss.losses import compute_sem_compl_loss
    torch.manual_seed(0)
    grid = (32, 32, 8)
    net = build_net(2, 64)
    net.class_weights = [torch.ones(K + 1) for _ in range(2)]
    net.criterion.class_weights = net.class_weights
    net.criterion.compl_labelweights = torch.ones(K)
    net.load_state_dict(fill_state_dict(net.state_dict()))
    net.train()
    batch = synthetic_batch(grid, 0.08, 2)
    g = torch.Generator().manual_seed(5)
    sem = torch.randint(0, K, (2,) + grid, generator=g)
    sem[torch.rand(sem.shape, generator=g) < 0.6] = 0
    sem[:, :, :2, :] = 255
    batch["semantic_label"] = sem.to(torch.uint8)
    batch["geo_labels"] = {"1_1": torch.where(sem == 255, 255.0, (sem > 0).float())}
    batch["mask_label"] = []
    for m in range(2):
        present = [c for c in range(1, K) if (sem[m] == c).any()]
        batch["mask_label"].append({"labels": torch.tensor(present, dtype=torch.uint8),
                                    "masks": torch.stack([sem[m] == c for c in present])})
    captured = {}
    crit_forward = net.criterion.forward

    def spy(*a, **k):
        out = crit_forward(*a, **k)
        captured.setdefault("crit", []).append(out)
        return out
    net.criterion.forward = spy
    net.evaluate_all = lambda *a, **k: None                  # metrics (CPU numpy, out of scope) — instance attribute only
    _rp = torch.randperm
    torch.randperm = lambda n, device=None, **k: _rp(n, **k)   # CylinderFeat asks get_device() == -1 on the CPU
    _fwd = net.forward

    def fwd_spy(*a, **k):
        captured["out"] = _fwd(*a, **k)
        return captured["out"]
    net.forward = fwd_spy
    try:
        ref = net.step(batch, "train")["loss"]
    finally:
        torch.randperm = _rp
    out = captured["out"]
    ce, lov = compute_sem_compl_loss(batch["sem_labels"], out["sem_logits_at_scales"], batch["min_Cs"], batch["max_Cs"],
                                     net.class_frequencies)
    total = (ce + lov) * net.occ_weight
    n = net.n_infers
    for o in captured["crit"]:
        total = total + (o["loss_dice"] + o["loss_ce"] + o["loss_mask"]) / n * net.panop_weight
        for v in o["loss_aux"].values():
            total = total + v / n
    print(f"Net.step loss {float(ref):.6f}   composed {float(total):.6f}")
    assert abs(float(ref) - float(total)) <= 1e-5 * abs(float(ref)), "composition differs from Net.step"
    assert isinstance(out["panop_predictions"][0]["voxel_logits"], ME.SparseTensor)


def main():
    torch.set_num_threads(os.cpu_count())
    if "--check-step" in sys.argv:
        check_against_net_step()
    d = synthetic()
    res, grads = reference_losses(d)
    arrays = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    arrays.update({k: np.asarray(v) for k, v in res.items()})
    arrays.update(grads)
    np.savez_compressed(os.path.join(HERE, "loss_m2.npz"), **arrays)
    print({k: (round(v, 6) if isinstance(v, float) else v.shape) for k, v in res.items()})
    print("bytes", os.path.getsize(os.path.join(HERE, "loss_m2.npz")))


if __name__ == "__main__":
    main()
