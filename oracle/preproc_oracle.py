"""CPU restatement of the reference's crop pre-processing  --  TEST INFRASTRUCTURE ONLY.

SURVEY.md section 8 row f1.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import
this; the product path (easy_vitpose_b200/csrc/preprocess.cuh) never does.

Parity: PINNED.  oracle/make_golden.py runs the UNMODIFIED reference (`pad_image`, `VitInference.pre_img`
and the whole `VitInference.inference` per-person loop, with a stub detector) on the same frames/boxes and
asserts that every function below reproduces it bit for bit (crops, org sizes, offsets) before it writes
tests/golden/frame_*.npz.

What the reference does per detected person (easy_ViTPose/inference.py):
  :253      bboxes = res_pd[:, :4].round().astype(int)
  :259-261  pad the box by 10 px and clip it to the frame
  :264      crop = img[y0:y1, x0:x1]
  :265      pad_image(crop, 3/4): zero-pad to a 3:4 (w:h) canvas             vit_utils/inference.py:41-70
  :314-318  pre_img: cv2.resize(.., (192, 256), INTER_LINEAR) on uint8, /255, (x - MEAN) / STD in float64,
            HWC -> 1xCxHxW, astype(float32)
  :270      keypoints[:, :2] += bbox[:2][::-1] - [top_pad, left_pad]

cv2's uint8 INTER_LINEAR is fixed-point (opencv modules/imgproc/src/resize.cpp, pinned here against
cv2 4.13 by experiment, 150 random sizes + degenerate ones, 0 differing pixels):
  * scale = 1 / (dst / src) in double; source position f = (float)((d + 0.5) * scale - 0.5); s = floor(f)
  * coefficients are int16: a1 = rint((f - s) * 2048), a0 = rint((1 - (f - s)) * 2048)
  * horizontally, s < 0 -> (s = 0, frac = 0) and s >= w - 1 -> (s = w - 1, frac = 0)
  * vertically the FRACTION is kept and only the two row indices are clamped to [0, h - 1]
  * rows: S = p[s] * a0 + p[s + 1] * a1 (int32);  columns: (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2
"""
from __future__ import annotations

import numpy as np

MEAN = (0.485, 0.456, 0.406)      # easy_ViTPose/inference.py:32
STD = (0.229, 0.224, 0.225)       # easy_ViTPose/inference.py:33
OUT_W, OUT_H = 192, 256           # data_cfg['image_size'], configs/ViTPose_common.py
PAD_BBOX = 10                     # easy_ViTPose/inference.py:255
COEF_BITS = 11                    # INTER_RESIZE_COEF_BITS


def normalise_lut() -> np.ndarray:
    """float32 [3, 256]: the value pre_img produces for channel c and byte v (float64 math, then cast)."""
    v = np.arange(256, dtype=np.uint8)[None, :] / 255                        # uint8 / int -> float64
    return ((v - np.array(MEAN)[:, None]) / np.array(STD)[:, None]).astype(np.float32)


def padded_box(bbox, frame_h: int, frame_w: int, pad: int = PAD_BBOX):
    """int box (x0, y0, x1, y1) -> the same after +-pad and clipping to the frame (inference.py:259-261)."""
    x0, y0, x1, y1 = (int(v) for v in bbox)
    x0 = min(max(x0 - pad, 0), frame_w); x1 = min(max(x1 + pad, 0), frame_w)
    y0 = min(max(y0 - pad, 0), frame_h); y1 = min(max(y1 + pad, 0), frame_h)
    return x0, y0, x1, y1


def pad_geometry(w: int, h: int):
    """Crop size -> (canvas_w, canvas_h, left_pad, top_pad) of pad_image(crop, 3/4) (vit_utils/inference.py:41-70).
    w / h < 3/4 is decided exactly as 4w < 3h; int(0.75 * h) = 3h // 4 and int(w / 0.75) = 4w // 3."""
    if w <= 0 or h <= 0:
        raise ValueError(f"empty crop {w}x{h}")                                # the reference divides by zero / cv2 asserts
    if 4 * w < 3 * h:
        cw = (3 * h) // 4
        return cw, h, (cw - w) // 2, 0
    ch = (4 * w) // 3
    return w, ch, 0, (ch - h) // 2


def _axis(dn: int, sn: int, clamp_fraction: bool):
    scale = 1.0 / (float(dn) / float(sn))
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    fr = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_fraction:
        lo, hi = s < 0, s >= sn - 1
        fr = np.where(lo | hi, np.float32(0), fr)
        s = np.where(lo, 0, np.where(hi, sn - 1, s))
    one = np.float32(1 << COEF_BITS)
    a1 = np.rint(fr * one).astype(np.int32)
    a0 = np.rint((np.float32(1) - fr) * one).astype(np.int32)
    i0 = np.clip(s, 0, sn - 1)
    i1 = np.clip(s + 1, 0, sn - 1)
    return i0, i1, a0, a1


def resize_linear_u8(src: np.ndarray, dw: int = OUT_W, dh: int = OUT_H) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 [h, w, c], bit-exact."""
    sh, sw = src.shape[:2]
    x0, x1, a0, a1 = _axis(dw, sw, True)
    y0, y1, b0, b1 = _axis(dh, sh, False)
    p = src.astype(np.int32)
    rows = p[:, x0] * a0[None, :, None] + p[:, x1] * a1[None, :, None]
    s0, s1 = rows[y0], rows[y1]
    out = (((b0[:, None, None] * (s0 >> 4)) >> 16) + ((b1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def pre_img(img: np.ndarray) -> np.ndarray:
    """uint8 RGB [h, w, 3] -> float32 [1, 3, 256, 192] (inference.py:314-318)."""
    r = resize_linear_u8(img)
    lut = normalise_lut()
    return np.stack([lut[c][r[..., c]] for c in range(3)], 0)[None]


def crop_canvas(frame: np.ndarray, bbox, pad: int = PAD_BBOX):
    """-> (zero-padded uint8 canvas [ch, cw, 3], (off_y, off_x)): the image handed to pre_img and the integer
    offset inference.py:270 adds to the crop-space keypoints."""
    fh, fw = frame.shape[:2]
    x0, y0, x1, y1 = padded_box(bbox, fh, fw, pad)
    w, h = x1 - x0, y1 - y0
    cw, ch, left, top = pad_geometry(w, h)
    canvas = np.zeros((ch, cw, 3), np.uint8)
    canvas[top:top + h, left:left + w] = frame[y0:y1, x0:x1]
    return canvas, (y0 - top, x0 - left)


def preprocess_frame(frame: np.ndarray, bboxes, pad: int = PAD_BBOX):
    """frame uint8 [H, W, 3] + int boxes [n, 4] -> crops f32 [n, 3, 256, 192], org_wh i32 [n, 2], offs_yx i32 [n, 2]."""
    crops, org, offs = [], [], []
    for b in np.asarray(bboxes).reshape(-1, 4):
        canvas, off = crop_canvas(frame, b, pad)
        crops.append(pre_img(canvas)[0])
        org.append((canvas.shape[1], canvas.shape[0]))
        offs.append(off)
    n = len(crops)
    return (np.stack(crops, 0) if n else np.zeros((0, 3, OUT_H, OUT_W), np.float32),
            np.asarray(org, np.int32).reshape(n, 2), np.asarray(offs, np.int32).reshape(n, 2))


def to_frame_coords(kpts: np.ndarray, offs_yx: np.ndarray) -> np.ndarray:
    """kpts f32 [n, K, 3] (y, x, score) in crop pixels -> frame pixels (inference.py:270; float32 result of
    an exact float64 sum, i.e. one rounding)."""
    out = kpts.copy()
    out[:, :, :2] = (kpts[:, :, :2].astype(np.float64) + offs_yx[:, None, :].astype(np.float64)).astype(np.float32)
    return out


def make_frame(h: int, w: int, seed: int) -> np.ndarray:
    """Synthetic uint8 RGB frame: smooth gradients and blobs (compressible) with bands of pixel noise so that
    every fixed-point rounding case of the resize is exercised."""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, 3))
    for c in range(3):
        img[..., c] = 96 + 70 * np.sin(xx / (17.0 + 5 * c) + c) + 60 * np.cos(yy / (23.0 - 4 * c))
        for _ in range(4):
            cy, cx, r = rs.uniform(0, h), rs.uniform(0, w), rs.uniform(6, 40)
            img[..., c] += rs.uniform(-90, 90) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * r * r))
    img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    band = (yy.astype(np.int64) // 16) % 3 == 0
    noise = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    img[band] = noise[band]
    return img
