"""-m gpu: SURVEY.md section 8 row f3 -- the reference's accuracy harness (evaluation_on_coco.py:31-87) run offline.

tests/golden/coco_ap.npz (oracle/make_golden_coco.py) holds a synthetic COCO-format set, the result records the UNMODIFIED
`VitInference.inference` loop produced on it with the reference's fp32 torch ViTPose behind a stub detector, and the ten
COCO keypoint summary numbers of those records.  Here the same frames and the same detector go through
`install(vi, batched=True)` (one engine call per frame: crop pre-processing, bf16 tensor-core model, decode, offsets on the
GPU); the records are built by the reference script's own rule (rounded pixel keypoints, detector score) and evaluated by
the same oracle (oracle/coco_oks_eval.py).  Bar: |AP_engine - AP_reference| <= 0.005 (and the same for AP50/75, AR)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import coco_oks_eval as E, preproc_oracle as P, vitpose_oracle as O

pytestmark = pytest.mark.gpu

AP_TOL = 0.005


def _ground_truth(g):
    return [{"id": i + 1, "image_id": int(g["gt_image"][i]), "category_id": 1, "iscrowd": 0, "num_keypoints": int(g["gt_num"][i]),
             "keypoints": g["gt_keypoints"][i].tolist(), "bbox": g["gt_bbox"][i].tolist(), "area": float(g["gt_area"][i])}
            for i in range(len(g["gt_image"]))]


def test_offline_coco_keypoint_ap_matches_the_reference(golden_dir):
    from easy_vitpose_b200 import install
    g = np.load(os.path.join(golden_dir, "coco_ap.npz"))
    n_img, fh, fw, fseed, D, depth, heads, K, wseed = (int(v) for v in g["meta"])
    sd = O.make_state_dict(D, depth, K, wseed, peaky=0.1, bumps=True)

    class FakeRefModel(torch.nn.Module):           # what install() needs of the reference ViTPose: state_dict() + num_heads
        def __init__(self):
            super().__init__()
            for k, v in sd.items():
                self.register_buffer(k.replace(".", "__"), torch.from_numpy(np.asarray(v)))
            self.backbone = types.SimpleNamespace(blocks=[types.SimpleNamespace(attn=types.SimpleNamespace(num_heads=heads))])

        def state_dict(self, *a, **kw):
            return {k.replace("__", "."): v for k, v in super().state_dict(*a, **kw).items()}

    rows_now = {}

    def yolo(img, **kw):
        data = types.SimpleNamespace(cpu=lambda: types.SimpleNamespace(numpy=lambda: rows_now["rows"]))
        return [types.SimpleNamespace(boxes=types.SimpleNamespace(data=data))]

    vi = types.SimpleNamespace(_vit_pose=FakeRefModel(), _inference=None, postprocess=None, tracker=None, frame_counter=0, yolo_step=1,
                               yolo=yolo, yolo_size=640, device="cuda", yolo_classes=[0], save_state=True)
    install(vi, max_batch=8, batched=True)

    results, kp_float = [], []
    image_ids = [1000 + i for i in range(n_img)]
    for i, image_id in enumerate(image_ids):
        rows_now["rows"] = g["rows"][i, :int(g["counts"][i])]
        frame = P.make_frame(fh, fw, fseed + i)
        out = vi.inference(frame)                                        # evaluation_on_coco.py:51
        results += E.results_from_frame_keypoints(image_id, out, vi._scores_bbox)      # :52-66
        kp_float += [np.asarray(out[k]) for k in sorted(out)]
    assert [r["image_id"] for r in results] == g["res_image"].tolist()
    assert np.allclose([r["score"] for r in results], g["res_score"])

    gts = _ground_truth(g)
    stats = E.evaluate(gts, results, image_ids)
    ref = dict(zip(g["stat_names"].tolist(), g["stat_values"].tolist()))
    # the stored reference numbers come from the stored reference records
    ref_records = [{"image_id": int(im), "category_id": 1, "score": float(s), "bbox": [], "keypoints": kp.tolist()}
                   for im, s, kp in zip(g["res_image"], g["res_score"], g["res_keypoints"])]
    again = E.evaluate(gts, ref_records, image_ids)
    assert all(abs(again[k] - ref[k]) < 1e-12 for k in ref)

    mine = np.array([r["keypoints"] for r in results]).reshape(len(results), K, 3)[..., :2]
    theirs = g["res_keypoints"].reshape(len(results), K, 3)[..., :2]
    same = float((mine == theirs).all(-1).mean())
    kpf = np.stack(kp_float, 0)
    vis = g["ref_kp"][..., 2] > 0.3
    dev = np.linalg.norm(kpf[..., :2] - g["ref_kp"][..., :2], axis=-1)
    print("offline COCO keypoint eval, engine vs reference:", {k: (round(stats[k], 4), round(ref[k], 4)) for k in ref})
    print(f"rounded pixel keypoints identical: {same:.4f}; float deviation over visible keypoints: median {np.median(dev[vis]):.4f} px, max {dev[vis].max():.3f} px")
    # AP is an average over 101 recall points x 10 OKS thresholds; AR moves in quanta of 1 / (people * 10) when one person
    # crosses one OKS threshold because a keypoint rounded to the neighbouring pixel: allow one quantum on top
    quantum = 1.0 / (len(gts) * 10)
    for k in ("AP", "AP50", "AP75", "AP_medium", "AP_large"):
        assert abs(stats[k] - ref[k]) <= AP_TOL, (k, stats[k], ref[k])
    for k in ("AR", "AR_medium", "AR_large"):
        assert abs(stats[k] - ref[k]) <= AP_TOL + quantum + 1e-12, (k, stats[k], ref[k])
    assert same > 0.9
