"""Golden vectors for pasco_b200/ensemble.py — BUILD CONTAINER ONLY: runs the reference's OWN Ensembler
(pasco/models/ensembler.py:20-187), find_matching_indices_v2 (utils.py:153-198) and panoptic_inference
(helper.py:91-303) from /root/reference on the CPU (MinkowskiEngine = the oracle shim) on a synthetic M=3 prediction set
with one identity pose, one flip + translation and one rotation.

    python tests/golden/make_golden_ensemble.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "shims"), os.path.join(ROOT, "compat"), "/root/reference", ROOT,
                os.path.join(ROOT, "oracle")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

Q, K, M, N = 24, 20, 3, 6000


def synthetic_predictions(seed=0):
    """M subnets looking at the same blobs under different poses: mask logits peak inside 12 boxes."""
    from pasco.models.transform_utils import generate_transformation, transform
    g = torch.Generator().manual_seed(seed)
    Ts = [torch.eye(4), generate_transformation(rot=0.0, translation=(0.4, -0.6, 0.0), flip_dim=1),
          generate_transformation(rot=7.0, translation=(0.0, 0.2, 0.0))]
    base = torch.stack([torch.randint(40, 200, (N,), generator=g), torch.randint(40, 200, (N,), generator=g),
                        torch.randint(2, 30, (N,), generator=g)], 1)
    base = torch.unique(base, dim=0)
    centres = base[torch.randperm(base.shape[0], generator=g)[:12]].float()
    preds, sems = [], []
    for m in range(M):
        # the subnet's own frame = T applied to the reference frame (sample_scene looks reference cells up through T)
        c = transform(base.float(), Ts[m]).int()
        c = torch.unique(c, dim=0)
        back = transform(c.float(), torch.inverse(Ts[m])).float()        # where each row sits in the reference frame
        d = (back[:, None, :] - centres[None]).abs().amax(-1)            # [n, 12]
        logits = torch.full((c.shape[0], Q), -9.0)
        perm = torch.randperm(Q, generator=g)                            # every subnet numbers its queries differently
        for j in range(12):
            logits[:, perm[j]] = 7.0 - 0.35 * d[:, j] + 0.3 * torch.randn(c.shape[0], generator=g)
        ql = torch.randn(1, Q, K + 1, generator=g)
        for j in range(12):
            ql[0, perm[j], 1 + (j % (K - 1))] += 4.0
        sem = torch.randn(c.shape[0], K, generator=g)
        sem[:, 1 + (torch.arange(c.shape[0]) % (K - 1))] += 2.0
        preds.append({"voxel_logits": (logits, c), "query_logits": ql})
        sems.append((sem, c))
    return preds, sems, Ts


def main():
    import MinkowskiEngine as ME
    from pasco.models.ensembler import Ensembler
    from pasco.models.helper import panoptic_inference
    preds, sems, Ts = synthetic_predictions()
    bc = lambda c: ME.utils.batched_coordinates([c])   # noqa: E731
    ens = Ensembler()
    sem_logits_at_scales = {1: [ME.SparseTensor(f, bc(c)) for f, c in sems]}
    sem_denses = ens.ensemble_sem_compl(sem_logits_at_scales, Ts)
    ref_preds = [{"voxel_logits": ME.SparseTensor(p["voxel_logits"][0], bc(p["voxel_logits"][1])),
                  "query_logits": p["query_logits"]} for p in preds]
    out = ens.ensemble_panop(ref_preds, sem_denses, (256, 256, 32), Ts, iou_threshold=0.2)
    arrays = {}
    for m in range(M):                      # the inputs travel with the fixture (the generator needs the reference)
        arrays[f"in_vox_logits{m}"], arrays[f"in_vox_coords{m}"] = preds[m]["voxel_logits"][0].numpy(), preds[m]["voxel_logits"][1].numpy()
        arrays[f"in_query_logits{m}"] = preds[m]["query_logits"].numpy()
        arrays[f"in_sem_logits{m}"], arrays[f"in_sem_coords{m}"] = sems[m][0].numpy(), sems[m][1].numpy()
        arrays[f"in_T{m}"] = Ts[m].numpy()
    for i, o in enumerate(out):
        C = o["voxel_probs"].C[:, 1:].long()
        lin = (C[:, 0] * 256 + C[:, 1]) * 32 + C[:, 2]
        order = torch.argsort(lin)
        arrays[f"vox{i}_lin"] = lin[order].numpy()
        arrays[f"vox{i}_F"] = o["voxel_probs"].F[order].numpy()
        arrays[f"query{i}"] = o["query_probs"].numpy()
        arrays[f"semrows{i}"] = o["sem_probs"].F[order].numpy()
    for i, d in enumerate(sem_denses):
        flat = d.reshape(d.shape[0], -1)
        arrays[f"semdense{i}_sub"] = flat[:, ::997].numpy()
        arrays[f"semdense{i}_sum"] = np.array([flat.double().sum().item(), float((flat.argmax(0) != 0).sum())])
    # panoptic inference on the ensembled prediction
    e = out[-1]
    res = panoptic_inference(e["voxel_probs"], e["query_probs"], overlap_threshold=0.5, object_mask_threshold=0.2,
                             thing_ids=[1, 2, 3, 4, 5, 6, 7, 8], min_C=torch.tensor([0, 0, 0]), scene_size=(256, 256, 32),
                             input_query_logit=False, input_voxel_logit=False)
    C = e["voxel_probs"].C[:, 1:].long()
    lin = (C[:, 0] * 256 + C[:, 1]) * 32 + C[:, 2]
    order = torch.argsort(lin)
    arrays["pan_sparse"] = res["panoptic_seg_sparses"][0][order].numpy()
    arrays["sem_at_rows"] = res["semantic_seg_denses"][0][C[:, 0], C[:, 1], C[:, 2]][order].numpy()
    arrays["segments"] = np.array([[s["id"], int(s["isthing"]), s["category_id"], s["query_id"]] for s in res["segments_infos"][0]])
    np.savez_compressed(os.path.join(HERE, "ensemble_m3.npz"), **arrays)
    print({k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    main()
