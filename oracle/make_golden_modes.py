"""Generate tests/golden/decode_modes.npz from the UNMODIFIED reference keypoints_from_heatmaps  --  TEST INFRASTRUCTURE ONLY.

Run here:  python oracle/make_golden_modes.py      
Maps regenerate from the seed (vitpose_oracle.make_decode_maps); stored: centre/scale in both dtypes and the reference's
(preds, maxvals) for every mode x use_udp x dtype combination the reference accepts, the blurring modes again with kernel=17
(sigma = 3 configs), and target_type='CombinedTarget' (kernel 11 and 17; one reference call per crop, the only batch size its
index arithmetic accepts).  Also asserts here that the oracle's zero-padded blur equals `_gaussian_blur` bit for bit on every
pixel and that its reflect-101 blur equals cv2.GaussianBlur on whole maps for every odd kernel 11..35.
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
N, K, SEED = 4, 17, 401
COMBOS = [(None, False), ("default", False), ("unbiased", False), ("megvii", False), ("default", True), ("unbiased", True)]
COMBOS_K17 = [("unbiased", False), ("megvii", False), ("default", True)]
KC = 6                                                      # keypoints of the CombinedTarget fixture (18 maps per crop)


def main() -> None:
    import importlib
    ns = ref_import.load()
    tde = importlib.import_module("vit_utils.top_down_eval")
    maps = O.make_decode_maps(N, K, SEED)
    taps = O.gaussian_taps(11)
    blurred = tde._gaussian_blur(maps.copy(), 11)
    for n in range(N):
        for k in range(K):
            mine = M.gaussian_modulate(maps[n, k], taps)
            assert np.array_equal(mine, blurred[n, k], equal_nan=True), (n, k, float(np.nanmax(np.abs(mine - blurred[n, k]))))
    print("zero-padded blur + renormalisation: bit-exact vs _gaussian_blur")
    rs = np.random.RandomState(SEED + 1)
    center32 = np.stack([rs.uniform(50, 600, N), rs.uniform(50, 400, N)], 1).astype(np.float32)
    scale32 = np.stack([rs.uniform(60, 400, N), rs.uniform(80, 520, N)], 1).astype(np.float32)
    scale64 = np.stack([rs.randint(64, 513, N), rs.randint(64, 513, N)], 1).astype(np.int64)
    center64 = np.stack([rs.randint(0, 900, N), rs.randint(0, 700, N)], 1).astype(np.int64)
    out = {"meta": np.array([N, K, SEED], np.int64), "center32": center32, "scale32": scale32, "center64": center64, "scale64": scale64}
    for pp, udp in COMBOS:
        for tag, (c, s) in {"f32": (center32, scale32), "i64": (center64, scale64)}.items():
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                preds, maxvals = ns.keypoints_from_heatmaps(maps.copy(), c, s, unbiased=False, post_process=pp, kernel=11, use_udp=udp)
            key = f"{pp}_{'udp' if udp else 'std'}_{tag}"
            out[key + "_preds"] = preds.astype(np.float32)
            out[key + "_maxvals"] = maxvals.astype(np.float32)
            opreds, omax, _ = M.keypoints_from_heatmaps(maps, c, s, post_process=pp, use_udp=udp)
            ok_max = np.array_equal(omax, maxvals.astype(np.float32), equal_nan=True)
            dev = np.nanmax(np.abs(opreds - preds))
            print(key, "oracle maxvals equal:", ok_max, "preds max |diff|:", float(dev), "dtype", preds.dtype)
    # the same with kernel = 17 (the modes that blur), and the blur itself for every kernel size the engine accepts
    import cv2
    for ks in range(11, 36, 2):
        tp = O.gaussian_taps(ks)
        assert np.array_equal(tp, cv2.getGaussianKernel(ks, 0).astype(np.float32).reshape(-1)), ks
        for n, k in ((0, 0), (1, 3), (2, 7), (3, 16)):
            assert np.array_equal(M.blur_reflect101(maps[n, k], tp), cv2.GaussianBlur(maps[n, k], (ks, ks), 0), equal_nan=True), (ks, n, k)
    print("reflect-101 blur: bit-exact vs cv2.GaussianBlur on whole maps, kernels 11..35")
    blurred = tde._gaussian_blur(maps.copy(), 17)
    for n in range(N):
        for k in range(K):
            assert np.array_equal(M.gaussian_modulate(maps[n, k], O.gaussian_taps(17)), blurred[n, k], equal_nan=True), (n, k)
    for pp, udp in COMBOS_K17:
        for tag, (c, s) in {"f32": (center32, scale32), "i64": (center64, scale64)}.items():
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                preds, maxvals = ns.keypoints_from_heatmaps(maps.copy(), c, s, unbiased=False, post_process=pp, kernel=17, use_udp=udp)
            key = f"k17_{pp}_{'udp' if udp else 'std'}_{tag}"
            out[key + "_preds"] = preds.astype(np.float32)
            out[key + "_maxvals"] = maxvals.astype(np.float32)
            opreds, omax, _ = M.keypoints_from_heatmaps(maps, c, s, post_process=pp, use_udp=udp, kernel=17)
            print(key, "oracle maxvals equal:", np.array_equal(omax, maxvals.astype(np.float32), equal_nan=True),
                  "preds max |diff|:", float(np.nanmax(np.abs(opreds - preds))))
    # CombinedTarget (:580-593): one call per crop
    cmaps = M.make_combined_maps(N, KC, SEED + 2)
    for ks in (11, 17):
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
                assert np.array_equal(o1[0], p1[0].astype(np.float32), equal_nan=True), (ks, tag, n, float(np.nanmax(np.abs(o1[0] - p1[0]))))
            key = f"comb_k{ks}_{tag}"
            out[key + "_preds"] = np.stack(pr).astype(np.float32)
            out[key + "_maxvals"] = np.stack(mv).astype(np.float32)
            print(key, "oracle == reference bit for bit; dtype", pr[0].dtype)
    try:
        ns.keypoints_from_heatmaps(cmaps.copy(), center32, scale32, kernel=11, use_udp=True, target_type="CombinedTarget")
        print("NOTE: the reference accepted N > 1 for CombinedTarget")
    except ValueError as e:
        print("CombinedTarget with N > 1 raises in the reference:", str(e)[:80])
    out["meta_combined"] = np.array([N, KC, SEED + 2], np.int64)
    # flip_back (post_transforms.py:110-147) + inference_model's shift (:210-212): data movement, pinned by digest
    import hashlib
    ptr = importlib.import_module("vit_utils.post_processing.post_transforms")
    for shift in (False, True):
        ref = ptr.flip_back(maps.copy(), M.COCO_FLIP_PAIRS, target_type="GaussianHeatmap")
        if shift:
            ref[:, :, :, 1:] = ref[:, :, :, :-1]
        mine = M.flip_back(maps, M.COCO_FLIP_PAIRS, shift)
        assert np.array_equal(mine, ref, equal_nan=True)
        out[f"flip_digest_shift{int(shift)}"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(ref).tobytes()).digest(), np.uint8)
    print("flip_back (+shift): oracle equals the reference")
    np.savez_compressed(os.path.join(OUT, "decode_modes.npz"), **out)
    print("written", os.path.getsize(os.path.join(OUT, "decode_modes.npz")), "bytes")


if __name__ == "__main__":
    main()
