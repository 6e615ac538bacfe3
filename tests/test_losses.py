"""Host-side checks of the training losses that close the benchmark's fwd+bwd loop (pasco_b200/losses.py).
The batched Lovász-softmax must equal the per-class loop of the published algorithm (Berman et al. 2018, as used by
pasco/loss/lovasz.py) in value and gradient."""
import torch
import torch.nn.functional as F

from pasco_b200 import losses


def _lovasz_loop(probs, labels):
    out = []
    for c in torch.unique(labels).tolist():
        fg = (labels == c).to(probs.dtype)
        err = (fg - probs[:, c]).abs()
        err_sorted, perm = torch.sort(err, descending=True)
        fg_sorted = fg[perm]
        gts = fg_sorted.sum()
        inter = gts - fg_sorted.cumsum(0)
        union = gts + (1 - fg_sorted).cumsum(0)
        jac = 1.0 - inter / union
        jac = torch.cat([jac[:1], jac[1:] - jac[:-1]])
        out.append(torch.dot(err_sorted, jac))
    return torch.stack(out).mean()


def test_batched_lovasz_equals_per_class_loop():
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(3000, 20, generator=g, dtype=torch.float64, requires_grad=True)
    labels = torch.randint(0, 13, (3000,), generator=g)            # classes 13..19 absent
    a = losses.lovasz_softmax_present(F.softmax(logits, 1), labels)
    b = _lovasz_loop(F.softmax(logits, 1), labels)
    ga, = torch.autograd.grad(a, logits)
    gb, = torch.autograd.grad(b, logits)
    assert abs(float(a) - float(b)) < 1e-12
    assert float((ga - gb).abs().max()) < 1e-12


def test_lovasz_empty_input_is_zero_with_grad():
    p = torch.zeros(0, 20, requires_grad=True)
    out = losses.lovasz_softmax_present(p, torch.zeros(0, dtype=torch.int64))
    assert float(out) == 0.0 and out.requires_grad


def test_masks_at_matches_box_membership():
    c = torch.tensor([[0, 1, 1, 1], [0, 9, 9, 9], [0, 10, 0, 0], [0, 5, 19, 10]], dtype=torch.int32)
    boxes = [((0, 0, 0), (10, 10, 10)), ((5, 5, 5), (30, 20, 11))]
    m = losses.masks_at(c, boxes)
    assert m.tolist() == [[1.0, 1.0, 0.0, 0.0], [0.0, 1.0, 0.0, 1.0]]


def test_oracle_copy_of_scene_and_losses_agrees_with_the_product():
    """oracle/scene_and_loss.py (the CPU arm's own copy, so that `bench.py --impl reference` imports nothing from the
    product) produces the same scene and the same loss pieces as pasco_b200.synthetic / pasco_b200.losses."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import scene_and_loss as OS
    from pasco_b200.synthetic import make_scene
    a, b = make_scene((32, 32, 8), 0.1, 1, seed=3), OS.make_scene((32, 32, 8), 0.1, 1, seed=3)
    assert torch.equal(a["in_coords"][0], b["in_coords"][0]) and torch.equal(a["in_feats"][0], b["in_feats"][0])
    assert all(torch.equal(a["sem_labels"][k], b["sem_labels"][k]) for k in a["sem_labels"]) and a["mask_boxes"] == b["mask_boxes"]
    g = torch.Generator().manual_seed(1)
    logits, labels = torch.randn(500, 20, generator=g, dtype=torch.float64), torch.randint(0, 20, (500,), generator=g)
    assert float(losses.lovasz_softmax_present(F.softmax(logits, 1), labels)) == float(OS.lovasz_softmax_present(F.softmax(logits, 1), labels))
    q, m = torch.randn(1, 30, 21, generator=g), torch.randn(400, 30, generator=g)
    tc, tm = torch.randint(1, 20, (6,), generator=g), (torch.rand(6, 400, generator=g) > 0.7).float()

    class _V:
        def __init__(self, f):
            self.F = f
    pred = {"query_logits": q, "voxel_logits": _V(m), "aux_outputs": []}
    assert abs(float(losses.panoptic_set_loss(pred, tc, tm, 20)) - float(OS.panoptic_set_loss(pred, tc, tm, 20))) < 1e-6
