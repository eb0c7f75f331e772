"""Generate the batch-size and outlier fixtures from the UNMODIFIED reference  --  TEST INFRASTRUCTURE ONLY.

Run here (the container that has /root/reference or baseline/_ref):   python oracle/make_golden_batch.py

Round-1 fixtures (make_golden.py) pin the path at B = 1..2.  BASELINE.json quotes the metric at B = 64 (ViT-B/17),
B = 32 (ViT-H/133) and 64 per GPU (ViT-L/25): these cases run the reference's ViTPose(cfg).forward + VitInference.postprocess
at exactly those batch sizes on seeded weights / crops and store
    kpts [B,K,3], idx [B,K] (np.argmax of the reference heatmaps), org_wh, range (min, max of all heatmaps),
    map_sum [B,K] float64 (a checksum of every map), and sample_hm [4 crops, 8 keypoints, 64, 48] with their indices.
Two more cases pin behaviour on real-ViT-like OUTLIERS (oracle.add_outliers): residual channels at +-100 and pre-GELU
activations at +-13 / +-26; full heatmaps are stored for those (B = 2).
Weights and inputs are regenerated from the seeds on both sides (np.random.RandomState is a frozen stream)."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, vitpose_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name -> (size, dataset, K, B, weight seed, crop seed, outlier seed or 0)
CASES = {
    "batch_b_coco_64": ("b", "coco", 17, 64, 111, 211, 0),
    "batch_h_wholebody_32": ("h", "wholebody", 133, 32, 112, 212, 0),
    "batch_l_coco_25_64": ("l", "coco_25", 25, 64, 113, 213, 0),
    "outlier_b_coco": ("b", "coco", 17, 2, 114, 214, 314),
    "outlier_l_coco_25": ("l", "coco_25", 25, 2, 115, 215, 315),
}


def org_sizes(B: int, seed: int) -> np.ndarray:
    rs = np.random.RandomState(seed + 7)
    return np.stack([rs.randint(64, 513, size=B), rs.randint(64, 513, size=B)], 1).astype(np.int32)


def main() -> None:
    import torch
    torch.set_grad_enabled(False)
    ns = ref_import.load()
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    for name, (size, dataset, K, B, wseed, xseed, oseed) in CASES.items():
        if only and name not in only:
            continue
        D, depth, heads = O.MODEL_DIMS[size]
        model = ns.ViTPose(ns.dyn_model_import(dataset, size)).eval()
        sd = O.make_state_dict(D, depth, K, wseed, peaky=0.1, bumps=True)
        if oseed:
            O.add_outliers(sd, oseed)
        model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        x = O.make_crops(B, xseed)
        hm = np.concatenate([model(torch.from_numpy(x[s:s + 16])).numpy() for s in range(0, B, 16)], 0).astype(np.float32)
        org_wh = org_sizes(B, xseed)
        kp = np.concatenate([ref_import.postprocess(ns, hm[i:i + 1], int(org_wh[i, 0]), int(org_wh[i, 1])) for i in range(B)], 0).astype(np.float32)
        idx = hm.reshape(B, K, -1).argmax(-1).astype(np.int32)
        rs = np.random.RandomState(xseed + 9)
        crop_ids = np.sort(rs.choice(B, size=min(4, B), replace=False)).astype(np.int32)
        kp_ids = np.sort(rs.choice(K, size=min(8, K), replace=False)).astype(np.int32)
        out = dict(kpts=kp, idx=idx, org_wh=org_wh, range=np.array([hm.min(), hm.max()], np.float32),
                   map_sum=hm.reshape(B, K, -1).sum(-1, dtype=np.float64), crop_ids=crop_ids, kp_ids=kp_ids,
                   sample_hm=hm[crop_ids][:, kp_ids], meta=np.array([D, depth, heads, K, B, wseed, xseed, oseed], np.int64))
        if oseed:
            out["heatmaps"] = hm
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
        vis = kp[..., 2] > 0.3
        print(name, "range", float(hm.min()), float(hm.max()), "visible (score > 0.3)", int(vis.sum()), "/", vis.size, flush=True)


if __name__ == "__main__":
    main()
