"""Generate tests/golden/decode_modes_small.npz from the UNMODIFIED reference keypoints_from_heatmaps  --  TEST INFRASTRUCTURE ONLY.

Run here:  python oracle/make_golden_modes_small.py      (modulation kernels below 11)

cv2.getGaussianKernel(k, sigma <= 0) returns fixed tables for k <= 9 and cv2's separable float filter sums kernels of 3 and 5
taps in a different order than wider ones (vitpose_oracle.row_pass).  This script pins both against cv2 itself on whole maps
(every odd kernel 1..35, reflect-101 and the reference's zero-padded `_gaussian_blur`) and stores the reference's outputs for
kernel = 1, 3, 5, 7, 9 in the modes that blur (kernel = 1 only with use_udp: `_gaussian_blur` raises for it), plus target_type='CombinedTarget' for kernel = 3 and 9 (response blurred with 7
and 19 taps).  Maps regenerate from the seed.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import decode_modes_oracle as M, ref_import, vitpose_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
N, K, SEED = 3, 17, 611
KERNELS = (1, 3, 5, 7, 9)
COMBOS = [("unbiased", False), ("megvii", False), ("default", True)]
KC = 5


def main() -> None:
    import importlib

    import cv2
    ns = ref_import.load()
    tde = importlib.import_module("vit_utils.top_down_eval")
    maps = O.make_decode_maps(N, K, SEED)
    for ks in range(1, 36, 2):
        tp = O.gaussian_taps(ks)
        assert np.array_equal(tp, cv2.getGaussianKernel(ks, 0).astype(np.float32).reshape(-1)), ks
        for n, k in ((0, 0), (1, 3), (2, 7), (2, 16)):
            assert np.array_equal(M.blur_reflect101(maps[n, k], tp), cv2.GaussianBlur(maps[n, k], (ks, ks), 0), equal_nan=True), (ks, n, k)
    print("taps and reflect-101 blur: bit-exact vs cv2 on whole maps, every odd kernel 1..35")
    try:
        tde._gaussian_blur(maps.copy(), 1)
        raise AssertionError("the reference accepted kernel = 1 in _gaussian_blur")
    except ValueError as e:                                      # border = 0: dr[0:-0] is empty
        print("kernel = 1 raises in the reference's _gaussian_blur (unbiased / megvii):", str(e)[:70])
    # `_gaussian_blur` blurs a zero-bordered copy that is 48 + 2r wide: for 5 and 7 taps the last visible columns fall into the
    # scalar (unfused) tail of cv2's column filter (decode_modes_oracle.blur_zero_padded)
    for ks in KERNELS[1:] + (11, 13, 17, 35):
        blurred = tde._gaussian_blur(maps.copy(), ks)
        for n in range(N):
            for k in range(K):
                mine = M.gaussian_modulate(maps[n, k], O.gaussian_taps(ks))
                assert np.array_equal(mine, blurred[n, k], equal_nan=True), (ks, n, k)
    print("zero-padded blur + renormalisation: bit-exact vs _gaussian_blur on every pixel, kernels", KERNELS[1:] + (11, 13, 17, 35))
    rs = np.random.RandomState(SEED + 1)
    center32 = np.stack([rs.uniform(50, 600, N), rs.uniform(50, 400, N)], 1).astype(np.float32)
    scale32 = np.stack([rs.uniform(60, 400, N), rs.uniform(80, 520, N)], 1).astype(np.float32)
    scale64 = np.stack([rs.randint(64, 513, N), rs.randint(64, 513, N)], 1).astype(np.int64)
    center64 = np.stack([rs.randint(0, 900, N), rs.randint(0, 700, N)], 1).astype(np.int64)
    out = {"meta": np.array([N, K, SEED], np.int64), "center32": center32, "scale32": scale32, "center64": center64, "scale64": scale64}
    for ks in KERNELS:
        for pp, udp in COMBOS:
            if ks == 1 and not udp:
                continue                                         # the reference raises (above)
            for tag, (c, s) in {"f32": (center32, scale32), "i64": (center64, scale64)}.items():
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    preds, maxvals = ns.keypoints_from_heatmaps(maps.copy(), c, s, unbiased=False, post_process=pp, kernel=ks, use_udp=udp)
                key = f"k{ks}_{pp}_{'udp' if udp else 'std'}_{tag}"
                out[key + "_preds"] = preds.astype(np.float32)
                out[key + "_maxvals"] = maxvals.astype(np.float32)
                opreds, omax, _ = M.keypoints_from_heatmaps(maps, c, s, post_process=pp, use_udp=udp, kernel=ks)
                assert np.array_equal(omax, maxvals.astype(np.float32), equal_nan=True), key
                dev = float(np.nanmax(np.abs(opreds - preds)))
                assert dev < (1e-6 if pp == "megvii" else 1e-3), (key, dev)
                print(key, "oracle maxvals equal, preds max |diff|:", dev)
    cmaps = M.make_combined_maps(N, KC, SEED + 2)
    for ks in (3, 9):
        for tag, (c, s) in {"f32": (center32, scale32), "i64": (center64, scale64)}.items():
            pr, mv = [], []
            for n in range(N):
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    p1, m1 = ns.keypoints_from_heatmaps(cmaps[n:n + 1].copy(), c[n:n + 1], s[n:n + 1], post_process="default", kernel=ks,
                                                        use_udp=True, target_type="CombinedTarget")
                pr.append(p1[0]); mv.append(m1[0])
                o1, om1, _ = M.combined_target(cmaps[n:n + 1], c[n:n + 1], s[n:n + 1], ks)
                assert np.array_equal(om1[0], m1[0].astype(np.float32), equal_nan=True), (ks, tag, n)
                assert np.array_equal(o1[0], p1[0].astype(np.float32), equal_nan=True), (ks, tag, n)
            out[f"comb_k{ks}_{tag}_preds"] = np.stack(pr).astype(np.float32)
            out[f"comb_k{ks}_{tag}_maxvals"] = np.stack(mv).astype(np.float32)
            print(f"comb_k{ks}_{tag}: oracle == reference bit for bit")
    out["meta_combined"] = np.array([N, KC, SEED + 2], np.int64)
    np.savez_compressed(os.path.join(OUT, "decode_modes_small.npz"), **out)
    print("written", os.path.getsize(os.path.join(OUT, "decode_modes_small.npz")), "bytes")


if __name__ == "__main__":
    main()
