"""Offline form of the reference's accuracy harness (evaluation_on_coco.py:31-87)  --  TEST INFRASTRUCTURE ONLY.

Run here (the container that has /root/reference or baseline/_ref):   python oracle/make_golden_coco.py

The real harness needs COCO val2017, a YOLO checkpoint, a ViTPose checkpoint and pycocotools; none can be mounted.
What CAN be kept is its flow, end to end, on a synthetic COCO-format set:
    for every image:  VitInference.inference(img)  ->  result records (keypoints rounded to pixels, score = detector
    confidence; evaluation_on_coco.py:52-66)  ->  COCO keypoint evaluation (OKS AP / AR; :69-81).
Here the UNMODIFIED `VitInference.inference` loop runs with the reference's torch ViTPose (fp32, CPU) behind a stub detector
that returns the annotated person boxes (jittered, with confidences), on seeded frames and seeded weights.  Ground-truth
keypoints are the reference's own float predictions perturbed by noise in OKS units (so that AP lands mid-range and every
OKS threshold from 0.5 to 0.95 discriminates), 15 % of them unlabelled.  The evaluation is oracle/coco_oks_eval.py.
Stored: detector rows, ground truth, the reference's result records and its ten summary numbers.  The -m gpu test
(tests/test_gpu_coco_ap.py) runs `install(vi, batched=True)` on the same frames and holds the engine's AP to the reference's.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import coco_oks_eval as E, preproc_oracle as P, ref_import, vitpose_oracle as O  # noqa: E402
from oracle.make_golden_frames import stub_detector  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
N_IMAGES, FH, FW, FSEED, SIZE, K, WSEED, GSEED = 24, 384, 512, 500, "b", 17, 121, 77


def detector_rows(rs) -> np.ndarray:
    n = int(rs.randint(1, 4))
    rows = []
    for _ in range(n):
        w = rs.uniform(60, 170); h = w * rs.uniform(1.7, 2.5)
        x0 = rs.uniform(-10, FW - w * 0.8); y0 = rs.uniform(-10, FH - h * 0.7)
        rows.append([x0, y0, x0 + w, y0 + h, rs.uniform(0.45, 0.97), 0.0])
    rows.append([30.0, 40.0, 90.0, 160.0, 0.2, 0.0])                       # below the 0.35 gate: never becomes a result
    return np.asarray(rows, np.float32)


def main() -> None:
    import torch
    torch.set_grad_enabled(False)
    ns = ref_import.load()
    VitInference = ref_import.load_vitinference().VitInference
    D, depth, heads = O.MODEL_DIMS[SIZE]
    model = ns.ViTPose(ns.dyn_model_import("coco", SIZE)).eval()
    sd = O.make_state_dict(D, depth, K, WSEED, peaky=0.1, bumps=True)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    rs = np.random.RandomState(GSEED)

    all_rows, results, gts, ref_kp = [], [], [], []
    for i in range(N_IMAGES):
        image_id = 1000 + i
        rows = detector_rows(rs)
        vi = object.__new__(VitInference)                  # __init__ loads YOLO / checkpoints from disk; set its fields by hand
        vi.tracker = None; vi.frame_counter = 0; vi.yolo_step = 1; vi.yolo_size = 640; vi.device = "cpu"
        vi.yolo_classes = [0]; vi.save_state = True; vi.dataset = "coco"; vi.target_size = (192, 256)
        vi.yolo = stub_detector(rows)
        vi._vit_pose = model
        vi._inference = lambda img, vi=vi: VitInference._inference_torch(vi, img)
        frame = P.make_frame(FH, FW, FSEED + i)
        out = vi.inference(frame)                          # evaluation_on_coco.py:51
        results += E.results_from_frame_keypoints(image_id, out, vi._scores_bbox)
        kept = rows[rows[:, 4] > 0.35]
        for pid, key in enumerate(sorted(out)):
            kp = np.asarray(out[key], np.float64)          # (y, x, score) in frame pixels
            x0, y0, x1, y1 = kept[pid, :4].astype(np.float64)
            bw, bh = x1 - x0, y1 - y0
            area = 0.6 * bw * bh
            s = rs.uniform(0.15, 0.85)                      # noise in OKS units: E[e] = s^2 per keypoint
            noise = rs.standard_normal((K, 2)) * (s * 2 * E.KPT_OKS_SIGMAS * np.sqrt(area))[:, None]
            v = np.where(rs.uniform(size=K) < 0.15, 0, 2)
            gk = np.stack([kp[:, 1] + noise[:, 0], kp[:, 0] + noise[:, 1], v], 1)
            gk[v == 0, :2] = 0
            gts.append({"id": len(gts) + 1, "image_id": image_id, "category_id": 1, "iscrowd": 0, "num_keypoints": int((v > 0).sum()),
                        "keypoints": gk.reshape(-1).tolist(), "bbox": [float(x0), float(y0), float(bw), float(bh)], "area": float(area)})
            ref_kp.append(kp)
        all_rows.append(rows)
    image_ids = [1000 + i for i in range(N_IMAGES)]
    stats = E.evaluate(gts, results, image_ids)
    print("people", len(gts), "reference stats", {k: round(v, 4) for k, v in stats.items()})
    nmax = max(len(r) for r in all_rows)
    rows_pad = np.zeros((N_IMAGES, nmax, 6), np.float32)
    counts = np.array([len(r) for r in all_rows], np.int32)
    for i, r in enumerate(all_rows):
        rows_pad[i, :len(r)] = r
    np.savez_compressed(
        os.path.join(OUT, "coco_ap.npz"), rows=rows_pad, counts=counts,
        gt_keypoints=np.array([g["keypoints"] for g in gts], np.float64), gt_bbox=np.array([g["bbox"] for g in gts], np.float64),
        gt_area=np.array([g["area"] for g in gts], np.float64), gt_image=np.array([g["image_id"] for g in gts], np.int64),
        gt_num=np.array([g["num_keypoints"] for g in gts], np.int64),
        res_keypoints=np.array([r["keypoints"] for r in results], np.float64), res_score=np.array([r["score"] for r in results], np.float64),
        res_image=np.array([r["image_id"] for r in results], np.int64), ref_kp=np.array(ref_kp, np.float32),
        stat_names=np.array(list(stats.keys())), stat_values=np.array(list(stats.values()), np.float64),
        meta=np.array([N_IMAGES, FH, FW, FSEED, D, depth, heads, K, WSEED], np.int64))
    print("written", os.path.getsize(os.path.join(OUT, "coco_ap.npz")), "bytes")


if __name__ == "__main__":
    main()
