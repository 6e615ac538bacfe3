"""pasco_b200/ensemble.py (sparse-row restatement of the reference's Ensembler + panoptic_inference) against fixtures the
reference's OWN functions produced (tests/golden/make_golden_ensemble.py → ensemble_m3.npz): M=3 subnets under an
identity, a flip+translation and a rotation pose.  Pure torch → runs on the CPU here and on the GPU unchanged."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
G = np.load(os.path.join(HERE, "golden", "ensemble_m3.npz"))
M = 3


def _inputs(dev):
    t = lambda a: torch.as_tensor(a).to(dev)  # noqa: E731
    preds = [{"voxel_logits": (t(G[f"in_vox_logits{m}"]), t(G[f"in_vox_coords{m}"])), "query_logits": t(G[f"in_query_logits{m}"])}
             for m in range(M)]
    sems = [(t(G[f"in_sem_logits{m}"]), t(G[f"in_sem_coords{m}"])) for m in range(M)]
    Ts = [torch.as_tensor(G[f"in_T{m}"]) for m in range(M)]
    return preds, sems, Ts


def _run(dev):
    from pasco_b200 import ensemble as E
    preds, sems, Ts = _inputs(dev)
    sem_denses = E.ensemble_sem_compl(sems, Ts)
    out = E.ensemble_panop(preds, sem_denses, Ts, iou_threshold=0.2)
    return E, sem_denses, out


def _check(dev):
    E, sem_denses, out = _run(dev)
    for i, d in enumerate(sem_denses):
        flat = d.reshape(d.shape[0], -1).cpu()
        assert np.allclose(flat[:, ::997].numpy(), G[f"semdense{i}_sub"], atol=1e-6), i
        assert int((flat.argmax(0) != 0).sum()) == int(G[f"semdense{i}_sum"][1])            # occupied cells: bit-exact
        assert abs(float(flat.double().sum()) - G[f"semdense{i}_sum"][0]) <= 1e-6 * G[f"semdense{i}_sum"][0]
    for i, o in enumerate(out):
        prob, lin = o["voxel_probs"]
        order = torch.argsort(lin)
        assert np.array_equal(lin[order].cpu().numpy(), G[f"vox{i}_lin"]), f"voxel set of output {i}"    # bit-exact rows
        assert prob.shape[1] == G[f"vox{i}_F"].shape[1], f"kept queries of output {i}"
        assert np.allclose(prob[order].cpu().numpy(), G[f"vox{i}_F"], atol=2e-6), i
        assert np.allclose(o["query_probs"].cpu().numpy(), G[f"query{i}"], atol=2e-6), i
        assert np.allclose(o["sem_probs"][order].cpu().numpy(), G[f"semrows{i}"], atol=2e-6), i
    prob, lin = out[-1]["voxel_probs"]
    pan, sem, infos = E.panoptic_inference(prob, out[-1]["query_probs"][0], overlap_threshold=0.5, object_mask_threshold=0.2,
                                           thing_ids=[1, 2, 3, 4, 5, 6, 7, 8])
    order = torch.argsort(lin)
    assert np.array_equal(pan[order].cpu().numpy(), G["pan_sparse"])
    assert np.array_equal(sem[order].cpu().numpy(), G["sem_at_rows"])
    got = np.array([[s["id"], int(s["isthing"]), s["category_id"], s["query_id"]] for s in infos])
    assert np.array_equal(got, G["segments"]) and len(infos) >= 4


def test_ensembler_and_panoptic_inference_match_the_reference_cpu():
    _check(torch.device("cpu"))


@pytest.mark.gpu
def test_ensembler_and_panoptic_inference_match_the_reference_gpu():
    _check(torch.device("cuda"))


def test_panoptic_inference_edge_cases():
    from pasco_b200 import ensemble as E
    q = torch.zeros(5, 21)
    q[:, 20] = 1.0                                                    # every query is "no object"
    pan, sem, infos = E.panoptic_inference(torch.rand(50, 5), q, 0.5, 0.2, [1, 2])
    assert int(pan.abs().sum()) == 0 and infos == []
    pan, sem, infos = E.panoptic_inference(torch.zeros(0, 5), torch.softmax(torch.randn(5, 21), -1), 0.5, 0.0, [1, 2])
    assert pan.numel() == 0 and infos == []
    # two queries of the same stuff class merge into one segment id; semantic labels of the merged part stay 0
    # (the reference `continue`s before writing semantic_seg, helper.py:243-245)
    q = torch.full((2, 21), 1e-3)
    q[:, 12] = 0.9
    v = torch.zeros(40, 2)
    v[:20, 0], v[20:, 1] = 0.9, 0.8
    pan, sem, infos = E.panoptic_inference(v, q, 0.5, 0.2, thing_ids=[1, 2, 3])
    assert set(pan.tolist()) == {1} and len(infos) == 1 and sem[:20].eq(12).all() and sem[20:].eq(0).all()
