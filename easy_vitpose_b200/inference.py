"""Drop-in boundary for easy_ViTPose/inference.py: the torch engine backend re-bound to the B200 engine.

The reference picks its engine in VitInference.__init__ by assigning `self._vit_pose` and
`self._inference = self._inference_torch` (easy_ViTPose/inference.py:156-172); the per-person loop then
calls `self._inference(img_inf)[0]` (:268).  `install()` performs the same two assignments with this
module's backend, so YOLO / SORT / draw and the CLI stay the reference's own code.

`B200PoseBackend` also offers what the reference lists as a TODO (README.md:323): one batched call for
all crops of a frame (`infer_crops`, `inference_batch`), and -- SURVEY.md section 8 rows f1/f2 -- the whole
per-person loop of `VitInference.inference` (:258-272) as ONE engine call on the uint8 frame
(`inference_frame`, and `install(..., batched=True)` which re-binds `VitInference.inference` itself).
"""
from __future__ import annotations

import types

import numpy as np
import torch

from .configs import data_cfg, model_cfg
from .model import ViTPose
from .top_down_eval import decode_heatmaps

__all__ = ["B200PoseBackend", "install", "frame_inference", "MEAN", "STD"]

MEAN = [0.485, 0.456, 0.406]      # easy_ViTPose/inference.py:32
STD = [0.229, 0.224, 0.225]       # easy_ViTPose/inference.py:33


def pre_img(img: np.ndarray, target_size=(192, 256)):
    """uint8 RGB [h,w,3] -> float32 [1,3,256,192] + (org_h, org_w): bilinear resize, /255, normalise in
    float64, HWC->CHW (easy_ViTPose/inference.py:314-318).  CPU, as in the reference."""
    import cv2
    org_h, org_w = img.shape[:2]
    x = cv2.resize(img, tuple(target_size), interpolation=cv2.INTER_LINEAR) / 255
    x = ((x - MEAN) / STD).transpose(2, 0, 1)[None].astype(np.float32)
    return x, org_h, org_w


class B200PoseBackend:
    """Owns a B200 `ViTPose` engine and exposes the three methods VitInference's torch backend consists of:
    pre_img (:314-318), _inference (:320-328), postprocess (:187-205)."""

    def __init__(self, model: ViTPose, device: "int | str | None" = None):
        self.model = model.eval()
        if device is not None:
            self.model.to(device)
        self.target_size = data_cfg["image_size"]

    @classmethod
    def from_state_dict(cls, state_dict: dict, size: str, num_keypoints: int, max_batch: int = 64, device=None):
        m = ViTPose(model_cfg(size, num_keypoints), max_batch=max_batch)
        m.load_state_dict(state_dict)
        return cls(m, device if device is not None else "cuda")

    def pre_img(self, img):
        return pre_img(img, self.target_size)

    @staticmethod
    def postprocess(heatmaps, org_w, org_h):
        """heatmaps [N,K,64,48] (numpy or CUDA tensor) -> float32 [N,K,3] rows (y, x, score).
        Same arguments as VitInference.postprocess; like the reference it treats the array as one call
        (for N=1 -- the only way VitInference uses it -- the two wrap modes coincide)."""
        hm = heatmaps if isinstance(heatmaps, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(heatmaps, np.float32))
        if not hm.is_cuda:
            hm = hm.cuda()
        n = hm.shape[0]
        org = torch.tensor([[int(org_w), int(org_h)]] * n, dtype=torch.int32)
        kp, _ = decode_heatmaps(hm, org, wrap_batch=True)
        return kp.cpu().numpy()

    @torch.no_grad()
    def _inference(self, img: np.ndarray) -> np.ndarray:
        """uint8 RGB crop -> float32 [1,K,3] (y, x, score) in crop pixels; the contract of
        VitInference._inference_torch (:320-328)."""
        x, org_h, org_w = self.pre_img(img)
        kp, _ = self.model.infer_host(x, np.array([[org_w, org_h]], np.int32))
        return kp

    @torch.no_grad()
    def infer_crops(self, crops, org_wh):
        """Pre-normalised crops [B,3,256,192] (CUDA tensor) -> (kpts [B,K,3], idx [B,K]) CUDA tensors."""
        return self.model.infer_crops(crops, org_wh)

    @torch.no_grad()
    def inference_frame(self, img: np.ndarray, bboxes: np.ndarray) -> np.ndarray:
        """uint8 RGB frame [H,W,3] + detector boxes [n,4] (x0,y0,x1,y1; floats are rounded as inference.py:253 does)
        -> float32 [n,K,3] (y, x, score) in frame pixels: inference.py:258-272 for all people in one engine call."""
        return self.model.infer_frame_host(img, bboxes)[0]

    @torch.no_grad()
    def inference_batch(self, imgs: "list[np.ndarray]") -> np.ndarray:
        """All person crops of a frame in one engine call -> float32 [n,K,3]."""
        if not imgs:
            return np.zeros((0, self.model.num_keypoints, 3), np.float32)
        pre = [self.pre_img(im) for im in imgs]
        x = np.concatenate([p[0] for p in pre], 0)
        org = np.array([[p[2], p[1]] for p in pre], np.int32)
        out = []
        for s in range(0, len(imgs), self.model.max_batch):
            kp, _ = self.model.infer_host(x[s:s + self.model.max_batch], org[s:s + self.model.max_batch])
            out.append(kp)
        return np.concatenate(out, 0)


def frame_inference(self, img: np.ndarray) -> dict:
    """`VitInference.inference` (easy_ViTPose/inference.py:214-281) with the per-person loop (:258-272) replaced by one
    engine call on the frame; bound onto the reference object by `install(..., batched=True)`, so `self` is the
    reference's VitInference: its detector (`self.yolo`), SORT tracker, counters and `save_state` fields are used and
    updated exactly as the reference does, which keeps `draw()` (:283-312) and the CLI's JSON writer working.

    Detection cadence (:234-241): the detector runs when there is no tracker, on the first three frames, and every
    `yolo_step`-th frame; rows with confidence <= 0.35 are dropped.  Returns {id: float32 [K,3] (y, x, score)} in frame pixels."""
    detections = np.empty((0, 5))
    results = None
    if self.tracker is None or self.frame_counter < 3 or self.frame_counter % self.yolo_step == 0:
        results = self.yolo(img[..., ::-1], verbose=False, imgsz=self.yolo_size,
                            device=0 if self.device == "cuda" else self.device, classes=self.yolo_classes)[0]
        rows = np.asarray(results.boxes.data.cpu().numpy(), np.float64)
        rows = rows.reshape(-1, rows.shape[-1] if rows.ndim > 1 else 6)
        detections = rows[rows[:, 4] > 0.35, :5].reshape(-1, 5)
    self.frame_counter += 1

    ids = None
    if self.tracker is not None:
        detections = self.tracker.update(detections)
        ids = detections[:, 5].astype(int).tolist()
    bboxes = detections[:, :4].round().astype(int)
    scores = detections[:, 4].tolist()
    if ids is None:
        ids = list(range(len(bboxes)))

    kpts, _ = self._b200.model.infer_frame_host(img, bboxes)            # pad/clip, crop, pad_image, pre_img, model, decode, offsets
    frame_keypoints = {i: kpts[n] for n, i in enumerate(ids)}
    scores_bbox = {i: sc for i, sc in zip(ids, scores)}

    if self.save_state:
        # the reference pads and clips `bboxes` in place inside its loop (:260-261), so draw() sees the padded boxes
        h, w = img.shape[:2]
        bboxes[:, [0, 2]] = np.clip(bboxes[:, [0, 2]] + [-10, 10], 0, w)
        bboxes[:, [1, 3]] = np.clip(bboxes[:, [1, 3]] + [-10, 10], 0, h)
        self._img = img
        self._yolo_res = results
        self._tracker_res = (bboxes, ids, scores)
        self._keypoints = frame_keypoints
        self._scores_bbox = scores_bbox
    return frame_keypoints


def install(vit_inference, max_batch: int = 64, device=None, batched: bool = False) -> B200PoseBackend:
    """Re-bind a constructed reference `VitInference` (torch .pth backend) to the B200 engine: takes the
    weights out of its `_vit_pose` module, then replaces `_vit_pose` and `_inference` exactly where
    easy_ViTPose/inference.py:156-172 set them.  With `batched=True` the object's `inference` method is re-bound to
    `frame_inference` as well (one engine call per frame instead of one per person).  Returns the backend (also stored
    as `._b200`)."""
    ref = vit_inference._vit_pose
    sd = {k: v.detach().cpu() for k, v in ref.state_dict().items()}
    D = sd["backbone.pos_embed"].shape[2]
    depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("backbone.blocks."))
    heads = int(ref.backbone.blocks[0].attn.num_heads)
    K = sd["keypoint_head.final_layer.weight"].shape[0]
    cfg = model_cfg({384: "s", 768: "b", 1024: "l", 1280: "h"}[D], K)
    cfg["backbone"].update(embed_dim=D, depth=depth, num_heads=heads)
    model = ViTPose(cfg, max_batch=max_batch)
    model.load_state_dict(sd)
    backend = B200PoseBackend(model, device if device is not None else "cuda")
    vit_inference._vit_pose = model
    vit_inference._inference = backend._inference
    vit_inference.postprocess = types.MethodType(lambda self, hm, w, h: backend.postprocess(hm, w, h), vit_inference)
    vit_inference._b200 = backend
    if batched:
        vit_inference.inference = types.MethodType(frame_inference, vit_inference)
    return backend
