"""Generate tests/golden/frame_*.npz from the UNMODIFIED reference  --  TEST INFRASTRUCTURE ONLY.

Run here (the container that has /root/reference):   python oracle/make_golden_frames.py

Pins SURVEY.md section 8 rows f1 / f2: the per-person loop of `VitInference.inference`
(easy_ViTPose/inference.py:232-281) is executed as it stands -- bbox rounding, +-10 px pad and clip, crop,
`pad_image`, `pre_img` (cv2 uint8 bilinear resize, float64 normalise), reference torch ViTPose on CPU,
`postprocess`, offset back to the frame -- with a stub detector that returns fixed boxes.  A spy on `_inference`
records what the loop hands to the pose engine.  Before anything is written, oracle/preproc_oracle.py is
asserted to reproduce the recorded canvases, the resized uint8 images, the float32 crops, the org sizes and
the frame offsets BIT FOR BIT.

Stored per case (frames and weights regenerate from seeds): the detector rows, the resized uint8 crops,
the normalise table, org_wh, offsets, and the reference's frame-space keypoints.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import preproc_oracle as P, ref_import, vitpose_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name -> (frame h, w, frame seed, model size, dataset, K, weight seed, detector rows [x0, y0, x1, y1, conf, cls])
FRAME_CASES = {
    "frame_a": (360, 480, 11, "s", "coco", 17, 101, [
        [200.4, 100.6, 290.5, 330.2, 0.91, 0],      # tall person: padded left/right
        [20.2, 40.7, 260.1, 140.3, 0.80, 0],        # wide box: padded top/bottom
        [-30.0, -12.0, 55.3, 170.8, 0.77, 0],       # clipped at the top-left corner
        [430.6, 250.2, 500.0, 380.0, 0.66, 0],      # clipped at the bottom-right corner
        [300.0, 200.0, 301.0, 201.0, 0.52, 0],      # 1x1 box -> 21x21 after padding: pure upscale
        [100.5, 101.5, 172.5, 217.5, 0.45, 0],      # .5 coordinates: round-half-even in the bbox cast
        [50.0, 60.0, 150.0, 260.0, 0.20, 0],        # below the 0.35 confidence gate: dropped
        [0.0, 0.0, 480.0, 360.0, 0.40, 0],          # the whole frame
    ]),
    "frame_b": (97, 131, 12, "s", "coco", 17, 101, [
        [10.0, 8.0, 60.0, 80.0, 0.9, 0],
        [70.2, 5.1, 128.9, 90.7, 0.8, 0],
        [40.0, 30.0, 43.0, 90.0, 0.7, 0],           # very thin box
    ]),
}


class _Boxes:
    def __init__(self, rows):
        self.data = self
        self._rows = np.asarray(rows, np.float32)

    def cpu(self):
        return self

    def numpy(self):
        return self._rows


def stub_detector(rows):
    """What `self.yolo(img, ...)` must look like to the loop: result[0].boxes.data.cpu().numpy() -> [n, 6]."""
    def yolo(img, **kwargs):
        return [types.SimpleNamespace(boxes=_Boxes(rows))]
    return yolo


def main() -> None:
    import cv2
    import torch
    torch.set_grad_enabled(False)
    ns = ref_import.load()
    inf = ref_import.load_vitinference()
    VitInference = inf.VitInference
    os.makedirs(OUT, exist_ok=True)
    lut = P.normalise_lut()

    for name, (fh, fw, fseed, size, dataset, K, wseed, rows) in FRAME_CASES.items():
        D, depth, heads = O.MODEL_DIMS[size]
        model = ns.ViTPose(ns.dyn_model_import(dataset, size)).eval()
        sd = O.make_state_dict(D, depth, K, wseed, peaky=0.1, bumps=True)
        model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)

        vi = object.__new__(VitInference)                  # __init__ loads YOLO / checkpoints from disk; set its fields by hand
        vi.tracker = None; vi.frame_counter = 0; vi.yolo_step = 1; vi.yolo_size = 320; vi.device = "cpu"
        vi.yolo_classes = [0]; vi.save_state = True; vi.dataset = dataset; vi.target_size = (192, 256)
        vi.yolo = stub_detector(rows)
        vi._vit_pose = model
        seen = []

        def spy(img, vi=vi, seen=seen):
            x, org_h, org_w = vi.pre_img(img)
            kp = VitInference._inference_torch(vi, img)
            seen.append((img.copy(), x.copy(), org_w, org_h, kp.copy()))
            return kp
        vi._inference = spy

        frame = P.make_frame(fh, fw, fseed)
        out = vi.inference(frame)
        ids = sorted(out.keys())
        n = len(ids)
        kp_frame = np.stack([out[i] for i in ids], 0).astype(np.float32)
        assert n == len(seen) == sum(1 for r in rows if r[4] > 0.35)

        # ---- the oracle against what the reference did, bit for bit
        kept = np.array([r for r in rows if r[4] > 0.35], np.float64)
        boxes = kept[:, :4].round().astype(int)
        crops, org_wh, offs = P.preprocess_frame(frame, boxes)
        resized = np.zeros((n, 256, 192, 3), np.uint8)
        for i, (img, x, org_w, org_h, kp) in enumerate(seen):
            canvas, off = P.crop_canvas(frame, boxes[i])
            assert np.array_equal(canvas, img), (name, i, "canvas")
            resized[i] = cv2.resize(img, (192, 256), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(P.resize_linear_u8(img), resized[i]), (name, i, "resize")
            assert np.array_equal(np.stack([lut[c][resized[i][..., c]] for c in range(3)], 0), x[0]), (name, i, "normalise")
            assert np.array_equal(crops[i], x[0]), (name, i, "crop")
            assert (org_wh[i, 0], org_wh[i, 1]) == (org_w, org_h), (name, i, "org")
            assert np.array_equal(P.to_frame_coords(kp, offs[i:i + 1])[0], kp_frame[i]), (name, i, "offset")
        # the numpy path oracle end to end (fp32, different summation order than torch: tolerance, not bits)
        kp_oracle = P.to_frame_coords(O.infer_crops(crops, org_wh, sd, depth, heads)[1], offs)
        vis = kp_frame[..., 2] > 0.3
        dev = np.abs(kp_oracle[..., :2] - kp_frame[..., :2])[vis]
        print(name, "n", n, "visible", int(vis.sum()), "/", vis.size, "oracle-vs-reference max px dev", float(dev.max()) if dev.size else None)

        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), rows=np.asarray(rows, np.float32), resized=resized, lut=lut,
                            org_wh=org_wh, offs_yx=offs, kpts=kp_frame,
                            meta=np.array([fh, fw, fseed, D, depth, heads, K, wseed], np.int64))
        print(name, "written", os.path.getsize(os.path.join(OUT, f"{name}.npz")), "bytes")


if __name__ == "__main__":
    main()
