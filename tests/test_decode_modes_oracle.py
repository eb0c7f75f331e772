"""CPU: oracle/decode_modes_oracle.py against the outputs of the UNMODIFIED reference
keypoints_from_heatmaps committed in tests/golden/decode_modes.npz (oracle/make_golden_modes.py)."""
import os

import numpy as np
import pytest

from oracle import decode_modes_oracle as M, vitpose_oracle as O

COMBOS = [(None, False), ("default", False), ("unbiased", False), ("megvii", False), ("default", True), ("unbiased", True)]


@pytest.fixture(scope="module")
def golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "decode_modes.npz"))
    N, K, seed = (int(v) for v in g["meta"])
    return g, O.make_decode_maps(N, K, seed)


@pytest.mark.parametrize("pp,udp", COMBOS)
@pytest.mark.parametrize("tag", ["f32", "i64"])
def test_modes_match_reference(golden, pp, udp, tag):
    g, maps = golden
    c, s = (g["center32"], g["scale32"]) if tag == "f32" else (g["center64"], g["scale64"])
    preds, maxvals, _ = M.keypoints_from_heatmaps(maps, c, s, post_process=pp, use_udp=udp)
    key = f"{pp}_{'udp' if udp else 'std'}_{tag}"
    assert np.array_equal(maxvals, g[key + "_maxvals"], equal_nan=True)
    ref = g[key + "_preds"]
    if pp in (None, "default", "megvii") and not udp:
        assert np.array_equal(preds, ref, equal_nan=True)          # integer / quarter-pixel arithmetic: bit-exact
    else:
        assert np.array_equal(np.isnan(preds), np.isnan(ref))
        assert np.nanmax(np.abs(preds - ref)) < 1e-3                # Taylor modes: float32 inverse


@pytest.mark.parametrize("pp,udp", [("unbiased", False), ("megvii", False), ("default", True)])
@pytest.mark.parametrize("tag", ["f32", "i64"])
def test_kernel_17_matches_reference(golden, pp, udp, tag):
    g, maps = golden
    c, s = (g["center32"], g["scale32"]) if tag == "f32" else (g["center64"], g["scale64"])
    preds, maxvals, _ = M.keypoints_from_heatmaps(maps, c, s, post_process=pp, use_udp=udp, kernel=17)
    key = f"k17_{pp}_{'udp' if udp else 'std'}_{tag}"
    assert np.array_equal(maxvals, g[key + "_maxvals"], equal_nan=True)
    ref = g[key + "_preds"]
    if pp == "megvii":
        assert np.array_equal(preds, ref, equal_nan=True)
    else:
        assert np.array_equal(np.isnan(preds), np.isnan(ref)) and np.nanmax(np.abs(preds - ref)) < 1e-3


@pytest.mark.parametrize("kernel", [11, 17])
@pytest.mark.parametrize("tag", ["f32", "i64"])
def test_combined_target_matches_reference(golden, kernel, tag):
    """One reference call per crop (the only batch size its index arithmetic accepts): bit-exact, sentinels included."""
    g, _ = golden
    N, KC, seed = (int(v) for v in g["meta_combined"])
    cmaps = M.make_combined_maps(N, KC, seed)
    c, s = (g["center32"], g["scale32"]) if tag == "f32" else (g["center64"], g["scale64"])
    sentinels = 0
    for n in range(N):
        preds, maxvals, _ = M.combined_target(cmaps[n:n + 1], c[n:n + 1], s[n:n + 1], kernel)
        assert np.array_equal(maxvals[0], g[f"comb_k{kernel}_{tag}_maxvals"][n], equal_nan=True)
        assert np.array_equal(preds[0], g[f"comb_k{kernel}_{tag}_preds"][n], equal_nan=True)
        sentinels += int((maxvals <= 0).sum())
    assert sentinels >= 2                                          # the fixture exercises the (-1,-1) offset lookup


def test_reflect_blur_is_blur_at_everywhere():
    """The whole-map reflect-101 blur and the point form the DARK decode uses are the same arithmetic."""
    maps = O.make_decode_maps(1, 3, 77)
    for ks in (11, 23, 35):
        taps = O.gaussian_taps(ks)
        full = M.blur_reflect101(maps[0, 1], taps)
        xs = np.array([0, 47, 5, 20, 47, 0]); ys = np.array([0, 63, 1, 30, 0, 63])
        assert np.array_equal(O.blur_at(maps[0, 1], xs, ys, taps), full[ys, xs])


def test_zero_padded_blur_properties():
    taps = O.gaussian_taps(11)
    h = np.zeros((64, 48), np.float32); h[30, 20] = 1.0
    g = M.blur_zero_padded(h, taps)
    assert np.array_equal(g[25:36, 15:26], np.outer(taps, taps).astype(np.float32))   # impulse response = the separable kernel
    corner = np.zeros((64, 48), np.float32); corner[0, 0] = 1.0
    gc = M.blur_zero_padded(corner, taps)
    assert gc[0, 0] == np.float32(taps[5] * taps[5]) and gc[6:, :].max() == 0            # zero border: nothing reflects back
    m = M.gaussian_modulate(h, taps)
    assert m.max() == np.float32(1.0)                                                   # maximum preserved (:456)


def test_flip_back_matches_reference_digest(golden):
    import hashlib
    g, maps = golden
    for shift in (False, True):
        mine = M.flip_back(maps, M.COCO_FLIP_PAIRS, shift)
        assert np.array_equal(np.frombuffer(hashlib.sha256(mine.tobytes()).digest(), np.uint8), g[f"flip_digest_shift{int(shift)}"])
    twice = M.flip_back(M.flip_back(maps, M.COCO_FLIP_PAIRS), M.COCO_FLIP_PAIRS)
    assert np.array_equal(twice, maps, equal_nan=True)                                 # an involution


@pytest.mark.parametrize("kernel", [1, 3, 5, 7, 9])
@pytest.mark.parametrize("tag", ["f32", "i64"])
def test_small_kernels_match_reference(golden_dir, kernel, tag):
    """Modulation kernels below 11 (tests/golden/decode_modes_small.npz, oracle/make_golden_modes_small.py): cv2 returns fixed tap
    tables there, sums 3 and 5 taps in its small-kernel order and runs the last visible columns of the zero-bordered 5- / 7-tap
    blur through the unfused tail of its column filter; the generating script pins all three on cv2 pixel by pixel."""
    g = np.load(os.path.join(golden_dir, "decode_modes_small.npz"))
    N, K, seed = (int(v) for v in g["meta"])
    maps = O.make_decode_maps(N, K, seed)
    c, s = (g["center32"], g["scale32"]) if tag == "f32" else (g["center64"], g["scale64"])
    for pp, udp in (("unbiased", False), ("megvii", False), ("default", True)):
        if kernel == 1 and not udp:
            continue                                                # the reference's _gaussian_blur raises for kernel = 1
        preds, maxvals, _ = M.keypoints_from_heatmaps(maps, c, s, post_process=pp, use_udp=udp, kernel=kernel)
        key = f"k{kernel}_{pp}_{'udp' if udp else 'std'}_{tag}"
        assert np.array_equal(maxvals, g[key + "_maxvals"], equal_nan=True)
        ref = g[key + "_preds"]
        if pp == "megvii":
            assert np.array_equal(preds, ref, equal_nan=True)
        else:
            assert np.array_equal(np.isnan(preds), np.isnan(ref)) and np.nanmax(np.abs(preds - ref)) < 1e-3


def test_small_kernel_taps_and_tail_rule():
    assert np.array_equal(O.gaussian_taps(9) * 256, [4, 13, 30, 51, 60, 51, 30, 13, 4])
    assert np.array_equal(O.gaussian_taps(5) * 16, [1, 4, 6, 4, 1]) and O.gaussian_taps(1).tolist() == [1.0]
    assert abs(float(O.gaussian_taps(11).sum()) - 1.0) < 1e-6
    # visible columns of a 48-wide map that cv2's column filter handles in its scalar tail when the map sits in a zero border
    assert [M.zero_padded_tail_start(48, r) for r in (1, 2, 3, 4, 5, 8, 17)] == [48, 46, 45, 48, 48, 48, 48]
