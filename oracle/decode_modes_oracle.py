"""CPU restatement of the OTHER decode modes of keypoints_from_heatmaps  --  TEST INFRASTRUCTURE ONLY.

SURVEY.md section 8 row f4.  The mode VitInference selects (unbiased=True, use_udp=True: DARK/UDP) lives in
oracle/vitpose_oracle.py::decode_maps; this file restates the remaining branches of
easy_ViTPose/vit_utils/top_down_eval.py:493-641 and the general (float centre / scale) form of transform_preds:

  post_process=None                :598 `_get_max_preds` only                                  (:82-114)
  post_process='default'           :617-631 +-0.25 px towards the higher neighbour
  post_process='unbiased'          :600-607 zero-padded Gaussian modulation (`_gaussian_blur`, :416-456),
                                   log, second-order Taylor step (`_taylor`, :315-350)
  post_process='megvii'            :573-574 blur first, then argmax of the BLURRED maps, +-0.25, +0.5, scores / 255 + 0.5
  use_udp=True (GaussianHeatmap)   :576-579 DARK/UDP with any centre / scale (decode_maps restricts centre to scale // 2)
  use_udp=True (CombinedTarget)    :580-593 response maps blurred with 2*kernel+1, offset maps with kernel, offsets at the arg-max
  kernel                           any odd modulation kernel 1..35 (cv2.getGaussianKernel computes 11 and up; <= 9 are fixed tables,
                                   3 and 5 taps take cv2's small-kernel row pass: vitpose_oracle.row_pass)
  transform_preds                  post_processing/post_transforms.py:150-194, both the /W (:186-187) and /(W-1) (:183-184) forms

Parity: PINNED.  oracle/make_golden_modes.py runs the unmodified reference function on seeded maps for every mode and both
centre/scale dtypes and stores its outputs in tests/golden/decode_modes.npz; tests/test_decode_modes_oracle.py holds this file
to them (argmax, scores, 'default'/'megvii'/None/CombinedTarget coordinates bit-exact; Taylor modes to 1e-3 px).

Arithmetic types follow numpy >= 2 (NEP 50: python scalars are weak), which is what the reference runs under here:
float32 centre/scale keep transform_preds in float32; int64 / float64 centre/scale promote it to float64.
"""
from __future__ import annotations

import numpy as np

from . import vitpose_oracle as O

MODES = (None, "default", "unbiased", "megvii")


def blur_zero_padded(h: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """`_gaussian_blur` for one map (:443-455) without the renormalisation: the map is embedded in a zero border as wide as the
    kernel radius, blurred with cv2.GaussianBlur and cropped, i.e. a zero-padded blur (cv2's own reflect border never reaches
    the cropped region).  Accumulation order = cv2's separable float filter (see vitpose_oracle.blur_at)."""
    H, W = h.shape
    r = (len(taps) - 1) // 2
    p = np.zeros((H + 2 * r, W + 2 * r), np.float32)
    p[r:r + H, r:r + W] = h
    rowpass = O.row_pass([p[:, j:j + W] for j in range(2 * r + 1)], taps)
    acc = (taps[r] * rowpass[r:r + H]).astype(np.float32)
    for d in range(1, r + 1):
        pair = (rowpass[r + d:r + d + H] + rowpass[r - d:r - d + H]).astype(np.float32)
        acc = O._fma32(np.broadcast_to(taps[r + d], acc.shape), pair, acc)
    # cv2's column filter is vectorised over x in steps of 8; the columns of the (W + 2r)-wide bordered image past the last
    # full step go through its scalar loop, whose products are NOT fused: acc += k[d] * (row[y+d] + row[y-d]) with two
    # roundings.  With W = 48 that reaches visible columns only for 5 and 7 taps (bordered widths 52 / 54: visible columns 46.. /
    # 45..; 3 taps go through cv2's small column filter, whose tail computes the same); found by matching cv2 bit for bit (oracle/make_golden_modes_small.py).
    tail = zero_padded_tail_start(W, r)
    if tail < W:
        t = (taps[r] * rowpass[r:r + H, tail:]).astype(np.float32)
        for d in range(1, r + 1):
            pair = (rowpass[r + d:r + d + H, tail:] + rowpass[r - d:r - d + H, tail:]).astype(np.float32)
            t = (t + (taps[r + d] * pair).astype(np.float32)).astype(np.float32)
        acc[:, tail:] = t
    return acc


def zero_padded_tail_start(W: int, r: int) -> int:
    """First visible column that cv2's column filter handles in its scalar (unfused) loop when the map is blurred inside a zero
    border of width r (`_gaussian_blur`, top_down_eval.py:443-455); W if none."""
    if r < 2:
        return W                                   # 3 taps: cv2's small column filter, the same arithmetic in its tail
    return min(W, max(0, 8 * ((W + 2 * r) // 8) - r))


def blur_reflect101(h: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """cv2.GaussianBlur(h, (k, k), 0) with its default BORDER_REFLECT_101 on the whole map (what CombinedTarget does in place,
    :582-584); vitpose_oracle.blur_at is the same thing at single points."""
    H, W = h.shape
    r = (len(taps) - 1) // 2
    p = np.pad(h.astype(np.float32), r, mode="reflect")
    rowpass = O.row_pass([p[:, j:j + W] for j in range(2 * r + 1)], taps)
    acc = (taps[r] * rowpass[r:r + H]).astype(np.float32)
    for d in range(1, r + 1):
        pair = (rowpass[r + d:r + d + H] + rowpass[r - d:r - d + H]).astype(np.float32)
        acc = O._fma32(np.broadcast_to(taps[r + d], acc.shape), pair, acc)
    return acc


def combined_target(heatmaps: np.ndarray, center, scale, kernel: int = 11, valid_radius_factor: float = 0.0546875):
    """use_udp=True, target_type='CombinedTarget' (:580-593): heatmaps [N,3K,H,W] -> (preds [N,K,2], maxvals [N,K,1], idx [N,K]).
    The reference's `index += W * H * np.arange(0, N * K / 3)` (:589) only broadcasts for N = 1; for N > 1 this is the same
    formula on the flattened [N*K] index (the shape the following reshape implies).  The flat index is formed in float32 like
    the reference's (`index` inherits preds' dtype) and the (-1,-1) sentinel lands W + 1 elements before the keypoint's plane,
    wrapping numpy-style for the very first one."""
    N, K3, H, W = heatmaps.shape
    if K3 % 3:
        raise ValueError("CombinedTarget needs triples of maps")
    K = K3 // 3
    wide, narrow = O.gaussian_taps(2 * kernel + 1), O.gaussian_taps(kernel)
    hm = np.empty((N, K3, H, W), np.float32)
    for n in range(N):
        for i in range(K3):
            hm[n, i] = blur_reflect101(heatmaps[n, i], wide if i % 3 == 0 else narrow)
    valid_radius = valid_radius_factor * H
    offset_x = (hm[:, 1::3].reshape(-1) * np.float32(valid_radius)).astype(np.float32)
    offset_y = (hm[:, 2::3].reshape(-1) * np.float32(valid_radius)).astype(np.float32)
    preds = np.empty((N, K, 2), np.float32)
    maxvals = np.empty((N, K, 1), np.float32)
    idxs = np.empty((N, K), np.int32)
    for n in range(N):
        for k in range(K):
            preds[n, k], maxvals[n, k, 0], idxs[n, k] = max_preds(hm[n, 3 * k])
    index = (preds[..., 0] + preds[..., 1] * np.float32(W)).astype(np.float32).reshape(-1)
    index = (index + (W * H * np.arange(0, N * K)).astype(np.float32)).astype(np.float32)
    index = index.astype(int).reshape(N, K, 1)
    preds = (preds + np.concatenate((offset_x[index], offset_y[index]), axis=2)).astype(np.float32)
    for n in range(N):
        preds[n] = transform(preds[n], center[n], scale[n], W, H, True)
    return preds, maxvals, idxs


def make_combined_maps(n: int, k: int, seed: int) -> np.ndarray:
    """[n,3k,64,48] CombinedTarget-style input: response maps from vitpose_oracle.make_decode_maps (blobs, sentinels, ties, ...)
    interleaved with smooth + noisy offset fields in about [-1, 1]."""
    resp = O.make_decode_maps(n, k, seed)
    rs = np.random.RandomState(seed + 7)
    H, W = resp.shape[2:]
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    out = np.empty((n, 3 * k, H, W), np.float32)
    out[:, 0::3] = resp
    for i in range(n):
        for j in range(k):
            for c in (1, 2):
                a, b, ph = rs.uniform(-0.2, 0.2, 3)
                out[i, 3 * j + c] = (np.sin(a * xx + b * yy + ph * 10) * rs.uniform(0.2, 1.0)
                                     + rs.standard_normal((H, W)) * 0.05).astype(np.float32)
    return out


def gaussian_modulate(h: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """One map of `_gaussian_blur` (:443-455): blur, then rescale so that the maximum is preserved."""
    origin_max = np.max(h)
    g = blur_zero_padded(h, taps)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (g * (origin_max / np.max(g))).astype(np.float32)


def max_preds(h: np.ndarray):
    """`_get_max_preds` for one map: ((x, y) float32 or (-1, -1), maxval, flat index)."""
    idx = int(np.argmax(h.reshape(-1)))
    mx = h.reshape(-1)[idx] if not np.isnan(h).any() else np.float32(np.nan)
    W = h.shape[1]
    if mx > 0.0:
        return np.array([idx % W, idx // W], np.float32), np.float32(mx), idx
    return np.array([-1, -1], np.float32), np.float32(mx), idx


def taylor(hm: np.ndarray, coord: np.ndarray) -> np.ndarray:
    """`_taylor` (:315-350) on a log-modulated map; float32 throughout, np.linalg.inv on a float32 2x2."""
    H, W = hm.shape
    px, py = int(coord[0]), int(coord[1])
    f = np.float32
    if 1 < px < W - 2 and 1 < py < H - 2:
        dx = f(0.5) * (hm[py][px + 1] - hm[py][px - 1])
        dy = f(0.5) * (hm[py + 1][px] - hm[py - 1][px])
        dxx = f(0.25) * (hm[py][px + 2] - f(2) * hm[py][px] + hm[py][px - 2])
        dxy = f(0.25) * (hm[py + 1][px + 1] - hm[py - 1][px + 1] - hm[py + 1][px - 1] + hm[py - 1][px - 1])
        dyy = f(0.25) * (hm[py + 2][px] - f(2) * hm[py][px] + hm[py - 2][px])
        if dxx * dyy - dxy ** 2 != 0:
            hinv = np.linalg.inv(np.array([[dxx, dxy], [dxy, dyy]], np.float32))
            off = -hinv @ np.array([[dx], [dy]], np.float32)
            coord = coord + off[:, 0]
    return coord.astype(np.float32)


def transform(coords: np.ndarray, center, scale, W: int, H: int, use_udp: bool) -> np.ndarray:
    """transform_preds for one crop's [K,2] coords.  numpy >= 2 promotion: float32 centre/scale -> float32 chain;
    anything else (int64, float64) -> float64 chain stored to float32."""
    center = np.asarray(center); scale = np.asarray(scale)
    if center.dtype == np.float32 and scale.dtype == np.float32:
        t = np.float32
    else:
        t = np.float64
        center = center.astype(np.float64); scale = scale.astype(np.float64)
    sx = t(scale[0] / t(W - 1.0 if use_udp else W)); sy = t(scale[1] / t(H - 1.0 if use_udp else H))
    out = np.empty_like(coords, dtype=np.float32)
    out[:, 0] = (coords[:, 0].astype(t) * sx + center[0] - scale[0] * t(0.5)).astype(np.float32)
    out[:, 1] = (coords[:, 1].astype(t) * sy + center[1] - scale[1] * t(0.5)).astype(np.float32)
    return out


def keypoints_from_heatmaps(heatmaps: np.ndarray, center: np.ndarray, scale: np.ndarray, post_process="default",
                            use_udp: bool = False, kernel: int = 11, target_type: str = "GaussianHeatmap",
                            valid_radius_factor: float = 0.0546875):
    """-> (preds [N,K,2] (x, y) float32, maxvals [N,K,1] float32, idx [N,K] int32 of the map the argmax was taken on)."""
    if use_udp and target_type.lower() == "combinedtarget":
        return combined_target(heatmaps, center, scale, kernel, valid_radius_factor)
    N, K, H, W = heatmaps.shape
    taps = O.gaussian_taps(kernel)
    preds = np.empty((N, K, 2), np.float32)
    maxvals = np.empty((N, K, 1), np.float32)
    idxs = np.empty((N, K), np.int32)
    if use_udp:
        if post_process == "megvii":
            raise AssertionError("use_udp excludes megvii (:557-558)")
        return _dark_udp_general(heatmaps, center, scale, kernel)
    for n in range(N):
        for k in range(K):
            h = heatmaps[n, k].astype(np.float32)
            if post_process == "megvii":
                h = gaussian_modulate(h, taps)
            c, mx, idx = max_preds(h)
            idxs[n, k] = idx
            if post_process == "unbiased":
                with np.errstate(divide="ignore", invalid="ignore"):
                    lg = np.log(np.maximum(gaussian_modulate(h, taps), np.float32(1e-10))).astype(np.float32)
                c = taylor(lg, c)
            elif post_process is not None:
                px, py = int(c[0]), int(c[1])
                if 1 < px < W - 1 and 1 < py < H - 1:
                    diff = np.array([h[py][px + 1] - h[py][px - 1], h[py + 1][px] - h[py - 1][px]], np.float32)
                    c = (c + np.sign(diff) * np.float32(0.25)).astype(np.float32)
                    if post_process == "megvii":
                        c = (c + np.float32(0.5)).astype(np.float32)
            preds[n, k] = c
            maxvals[n, k, 0] = mx
        preds[n] = transform(preds[n], center[n], scale[n], W, H, False)
    if post_process == "megvii":
        maxvals = (maxvals / np.float32(255.0) + np.float32(0.5)).astype(np.float32)
    return preds, maxvals, idxs


def _dark_udp_general(heatmaps, center, scale, kernel):
    """post_dark_udp (:354-415) followed by transform_preds with ARBITRARY centre / scale.  vitpose_oracle.decode_maps restates
    post_dark_udp but maps with integer org sizes and centre = org // 2; decoding with org = (2(W-1), 2(H-1)) makes that map
    X = x * 2.0 + (W-1) - (W-1) = 2x exactly, so the refined heatmap-pixel coordinates are recovered as X / 2 without any
    extra rounding and then transformed here."""
    N, K, H, W = heatmaps.shape
    org = np.tile(np.array([[2 * (W - 1), 2 * (H - 1)]], np.int32), (N, 1))
    kp, idx = O.decode_maps(heatmaps, org, wrap="batch", ksize=kernel)
    preds = np.empty((N, K, 2), np.float32)
    for n in range(N):
        xy = np.stack([kp[n, :, 1] * np.float32(0.5), kp[n, :, 0] * np.float32(0.5)], 1).astype(np.float32)
        preds[n] = transform(xy, center[n], scale[n], W, H, True)
    return preds, kp[:, :, 2:3].copy(), idx


COCO_FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]     # datasets/COCO.py:114


def flip_back(heatmaps: np.ndarray, flip_pairs, shift_heatmap: bool = False) -> np.ndarray:
    """flip_back for GaussianHeatmap (post_processing/post_transforms.py:110-147) followed by the optional one-pixel shift of
    TopdownHeatmapSimpleHead.inference_model (head/topdown_heatmap_simple_head.py:210-212)."""
    out = heatmaps.copy()
    for left, right in flip_pairs:
        out[:, left] = heatmaps[:, right]
        out[:, right] = heatmaps[:, left]
    out = np.ascontiguousarray(out[..., ::-1])
    if shift_heatmap:
        out[:, :, :, 1:] = out[:, :, :, :-1].copy()
    return out
