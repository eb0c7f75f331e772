"""-m gpu: every GaussianHeatmap branch of keypoints_from_heatmaps on the GPU against the
unmodified reference's outputs (tests/golden/decode_modes.npz) and against the oracle on more maps."""
import os

import numpy as np
import pytest
import torch

from oracle import decode_modes_oracle as M, vitpose_oracle as O

pytestmark = pytest.mark.gpu

COMBOS = [(None, False), ("default", False), ("unbiased", False), ("megvii", False), ("default", True), ("unbiased", True)]


def _check(preds, maxvals, ref_preds, ref_maxvals, exact):
    assert preds.dtype == np.float32 and maxvals.dtype == np.float32 and maxvals.shape == ref_maxvals.shape
    assert np.array_equal(maxvals, ref_maxvals, equal_nan=True)                       # scores: bit-exact in every mode
    assert np.array_equal(np.isnan(preds), np.isnan(ref_preds))
    if exact:
        assert np.array_equal(preds, ref_preds, equal_nan=True)                       # argmax / quarter-pixel modes: bit-exact
    else:
        # Taylor modes: logf vs np.log ulps through a 2x2 solve; image pixels (scales up to ~8 px per heatmap cell)
        assert np.nanmax(np.abs(preds - ref_preds)) < 2e-2


@pytest.mark.parametrize("pp,udp", COMBOS)
@pytest.mark.parametrize("tag", ["f32", "i64"])
def test_modes_vs_reference_fixture(golden_dir, pp, udp, tag):
    from easy_vitpose_b200 import keypoints_from_heatmaps
    g = np.load(os.path.join(golden_dir, "decode_modes.npz"))
    N, K, seed = (int(v) for v in g["meta"])
    maps = O.make_decode_maps(N, K, seed)
    c, s = (g["center32"], g["scale32"]) if tag == "f32" else (g["center64"], g["scale64"])
    before = maps.copy()
    preds, maxvals = keypoints_from_heatmaps(maps, c, s, post_process=pp, use_udp=udp)
    assert np.array_equal(maps, before, equal_nan=True)                               # the caller's heatmaps are not modified
    key = f"{pp}_{'udp' if udp else 'std'}_{tag}"
    _check(preds, maxvals, g[key + "_preds"], g[key + "_maxvals"], exact=pp in (None, "default", "megvii") and not udp)


@pytest.mark.parametrize("pp,udp", COMBOS)
def test_modes_vs_oracle_on_more_maps(pp, udp):
    from easy_vitpose_b200 import keypoints_from_heatmaps
    N, K = 5, 25
    maps = O.make_decode_maps(N, K, 977)
    rs = np.random.RandomState(3)
    c = rs.uniform(10, 800, (N, 2)).astype(np.float32); s = rs.uniform(40, 500, (N, 2)).astype(np.float32)
    preds, maxvals, idx = keypoints_from_heatmaps(torch.from_numpy(maps).cuda(), c, s, post_process=pp, use_udp=udp, return_idx=True)
    op, om, oi = M.keypoints_from_heatmaps(maps, c, s, post_process=pp, use_udp=udp)
    assert np.array_equal(idx, oi)                                                    # integer argmax (megvii: of the blurred map)
    _check(preds, maxvals, op, om, exact=pp in (None, "default", "megvii") and not udp)


@pytest.mark.parametrize("pp,udp", [("unbiased", False), ("megvii", False), ("default", True)])
@pytest.mark.parametrize("tag", ["f32", "i64"])
def test_kernel_17_vs_reference_fixture(golden_dir, pp, udp, tag):
    """modulate_kernel = 17 (sigma = 3 configs): taps computed on the host like cv2.getGaussianKernel, same accumulation order."""
    from easy_vitpose_b200 import keypoints_from_heatmaps
    g = np.load(os.path.join(golden_dir, "decode_modes.npz"))
    N, K, seed = (int(v) for v in g["meta"])
    maps = O.make_decode_maps(N, K, seed)
    c, s = (g["center32"], g["scale32"]) if tag == "f32" else (g["center64"], g["scale64"])
    preds, maxvals = keypoints_from_heatmaps(maps, c, s, post_process=pp, use_udp=udp, kernel=17)
    key = f"k17_{pp}_{'udp' if udp else 'std'}_{tag}"
    _check(preds, maxvals, g[key + "_preds"], g[key + "_maxvals"], exact=pp == "megvii")


@pytest.mark.parametrize("kernel", [1, 3, 5, 7, 9])
@pytest.mark.parametrize("tag", ["f32", "i64"])
def test_small_kernels_vs_reference_fixture(golden_dir, kernel, tag):
    """Modulation kernels below 11 against the unmodified reference (tests/golden/decode_modes_small.npz): cv2's fixed tap tables,
    its small-kernel row order for 3 / 5 taps and the unfused tail of its column filter for 5 / 7 taps in the zero-bordered blur.
    megvii (scores and quarter-pixel coordinates straight from the blurred map) is bit-exact; kernel = 1 only exists with use_udp."""
    from easy_vitpose_b200 import keypoints_from_heatmaps
    g = np.load(os.path.join(golden_dir, "decode_modes_small.npz"))
    N, K, seed = (int(v) for v in g["meta"])
    maps = O.make_decode_maps(N, K, seed)
    c, s = (g["center32"], g["scale32"]) if tag == "f32" else (g["center64"], g["scale64"])
    for pp, udp in (("unbiased", False), ("megvii", False), ("default", True)):
        if kernel == 1 and not udp:
            continue
        preds, maxvals = keypoints_from_heatmaps(maps, c, s, post_process=pp, use_udp=udp, kernel=kernel)
        key = f"k{kernel}_{pp}_{'udp' if udp else 'std'}_{tag}"
        _check(preds, maxvals, g[key + "_preds"], g[key + "_maxvals"], exact=pp == "megvii")


@pytest.mark.parametrize("kernel", [3, 5, 7, 9])
def test_small_kernels_vs_oracle_on_more_maps(kernel):
    """The whole blurred map is what megvii's arg-max runs over: index, score and coordinates bit-exact against the oracle (which
    is pinned on cv2 pixel by pixel) on maps whose peaks sit in the last columns too (the scalar tail of cv2's column filter)."""
    from easy_vitpose_b200 import keypoints_from_heatmaps
    N, K = 4, 12
    maps = O.make_decode_maps(N, K, 3100 + kernel)
    maps[:, :, :, 40:] += maps[:, :, :, 8:0:-1] * 1.5                                 # mass near the right border
    rs = np.random.RandomState(kernel)
    c = rs.uniform(10, 800, (N, 2)).astype(np.float32); s = rs.uniform(40, 500, (N, 2)).astype(np.float32)
    for pp, udp in (("megvii", False), ("unbiased", False), ("default", True)):
        preds, maxvals, idx = keypoints_from_heatmaps(maps, c, s, post_process=pp, use_udp=udp, kernel=kernel, return_idx=True)
        op, om, oi = M.keypoints_from_heatmaps(maps, c, s, post_process=pp, use_udp=udp, kernel=kernel)
        assert np.array_equal(idx, oi)
        _check(preds, maxvals, op, om, exact=pp == "megvii")


@pytest.mark.parametrize("kernel", [3, 9])
@pytest.mark.parametrize("tag", ["f32", "i64"])
def test_combined_target_small_kernels_vs_reference_fixture(golden_dir, kernel, tag):
    from easy_vitpose_b200 import keypoints_from_heatmaps
    g = np.load(os.path.join(golden_dir, "decode_modes_small.npz"))
    N, KC, seed = (int(v) for v in g["meta_combined"])
    cmaps = M.make_combined_maps(N, KC, seed)
    c, s = (g["center32"], g["scale32"]) if tag == "f32" else (g["center64"], g["scale64"])
    for n in range(N):
        preds, maxvals = keypoints_from_heatmaps(cmaps[n:n + 1], c[n:n + 1], s[n:n + 1], kernel=kernel, use_udp=True, target_type="CombinedTarget")
        assert np.array_equal(maxvals[0], g[f"comb_k{kernel}_{tag}_maxvals"][n], equal_nan=True)
        assert np.array_equal(preds[0], g[f"comb_k{kernel}_{tag}_preds"][n], equal_nan=True)


@pytest.mark.parametrize("kernel", [13, 23, 35])
def test_other_kernels_vs_oracle(kernel):
    """Blurred values are bit-exact for every kernel size: megvii's scores and quarter-pixel coordinates come straight from them."""
    from easy_vitpose_b200 import keypoints_from_heatmaps
    N, K = 3, 10
    maps = O.make_decode_maps(N, K, 1200 + kernel)
    rs = np.random.RandomState(kernel)
    c = rs.uniform(10, 800, (N, 2)).astype(np.float32); s = rs.uniform(40, 500, (N, 2)).astype(np.float32)
    for pp, udp in (("megvii", False), ("unbiased", False), ("default", True)):
        preds, maxvals, idx = keypoints_from_heatmaps(maps, c, s, post_process=pp, use_udp=udp, kernel=kernel, return_idx=True)
        op, om, oi = M.keypoints_from_heatmaps(maps, c, s, post_process=pp, use_udp=udp, kernel=kernel)
        assert np.array_equal(idx, oi)
        _check(preds, maxvals, op, om, exact=pp == "megvii")


@pytest.mark.parametrize("kernel", [11, 17])
@pytest.mark.parametrize("tag", ["f32", "i64"])
def test_combined_target_vs_reference_fixture(golden_dir, kernel, tag):
    """use_udp=True, target_type='CombinedTarget' (top_down_eval.py:580-593), one call per crop like the fixture: blurred response
    maximum, its index and the blurred offsets are bit-exact, so the keypoints are too."""
    from easy_vitpose_b200 import keypoints_from_heatmaps
    g = np.load(os.path.join(golden_dir, "decode_modes.npz"))
    N, KC, seed = (int(v) for v in g["meta_combined"])
    cmaps = M.make_combined_maps(N, KC, seed)
    c, s = (g["center32"], g["scale32"]) if tag == "f32" else (g["center64"], g["scale64"])
    for n in range(N):
        preds, maxvals = keypoints_from_heatmaps(cmaps[n:n + 1], c[n:n + 1], s[n:n + 1], kernel=kernel, use_udp=True, target_type="CombinedTarget")
        assert preds.shape == (1, KC, 2) and maxvals.shape == (1, KC, 1)
        assert np.array_equal(maxvals[0], g[f"comb_k{kernel}_{tag}_maxvals"][n], equal_nan=True)
        assert np.array_equal(preds[0], g[f"comb_k{kernel}_{tag}_preds"][n], equal_nan=True)


def test_combined_target_batched_through_the_c_abi():
    """The C ABI takes any n for mode 5 (the reference's formula on the flattened index): against the oracle, including the
    sentinel's wrap to the last plane of the call."""
    import ctypes as C
    from easy_vitpose_b200 import _lib
    N, KC = 3, 7
    cmaps = M.make_combined_maps(N, KC, 4242)
    cmaps[0, 0] = -np.abs(cmaps[0, 0]) - 0.1                                           # first keypoint of the call: sentinel
    rs = np.random.RandomState(8)
    c = rs.uniform(10, 800, (N, 2)).astype(np.float32); s = rs.uniform(40, 500, (N, 2)).astype(np.float32)
    hm = torch.from_numpy(cmaps).cuda()
    cs = torch.from_numpy(np.concatenate([c, s], 1)).cuda()
    kp = torch.empty((N, KC, 3), dtype=torch.float32, device="cuda"); idx = torch.empty((N, KC), dtype=torch.int32, device="cuda")
    for kernel in (11, 17):
        _lib.check(_lib.lib().vpb_decode_modes_ex(C.c_void_p(hm.data_ptr()), N, KC, 5, kernel, float(np.float32(0.0546875 * 64)),
                                                  C.c_void_p(cs.data_ptr()), None, C.c_void_p(kp.data_ptr()), C.c_void_p(idx.data_ptr()),
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        op, om, oi = M.combined_target(cmaps, c, s, kernel)
        k = kp.cpu().numpy()
        assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(k[..., 2:3], om, equal_nan=True)
        assert np.array_equal(k[..., 1::-1], op, equal_nan=True)
    assert _lib.lib().vpb_decode_modes_ex(C.c_void_p(hm.data_ptr()), N, KC, 5, 19, 3.5, C.c_void_p(cs.data_ptr()), None,
                                          C.c_void_p(kp.data_ptr()), None, None) != 0   # 2 * 19 + 1 > 35
    assert _lib.lib().vpb_decode_modes_ex(C.c_void_p(hm.data_ptr()), N, KC, 2, 12, 0.0, C.c_void_p(cs.data_ptr()), None,
                                          C.c_void_p(kp.data_ptr()), None, None) != 0   # even kernel


def test_config_normalisation_and_errors():
    from easy_vitpose_b200 import decode_topdown, keypoints_from_heatmaps
    maps = O.make_decode_maps(2, 17, 5)
    c = np.array([[96, 128], [100, 100]], np.int64); s = np.array([[192, 256], [200, 201]], np.int64)
    a = keypoints_from_heatmaps(maps, c, s, unbiased=True, post_process="default", use_udp=True)      # VitInference's call
    b = keypoints_from_heatmaps(maps, c, s, post_process="unbiased", use_udp=True)
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))
    d = keypoints_from_heatmaps(maps, c, s, unbiased=True, post_process=True)                           # deprecated spellings
    e = keypoints_from_heatmaps(maps, c, s, post_process="unbiased")
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(d, e))
    with pytest.raises(AssertionError):
        keypoints_from_heatmaps(maps, c, s, post_process="megvii", use_udp=True)
    with pytest.raises(ValueError):
        keypoints_from_heatmaps(maps, c, s, use_udp=True, target_type="CombinedTarget")   # N = 2, 17 maps: as in the reference
    with pytest.raises(NotImplementedError):
        keypoints_from_heatmaps(maps, c, s, post_process="unbiased", kernel=10)
    with pytest.raises(ValueError):                                                       # kernel = 1: the reference's _gaussian_blur raises
        keypoints_from_heatmaps(maps, c, s, post_process="megvii", kernel=1)
    # TopdownHeatmapBaseHead.decode with the reference's test_cfg (configs/ViTPose_common.py:123-129)
    metas = [{"center": [96.5, 128.0], "scale": [192.0, 256.0], "image_file": "a.jpg", "bbox_score": 0.9, "bbox_id": 7},
             {"center": [50.0, 60.0], "scale": [120.0, 160.0], "image_file": "b.jpg", "bbox_id": 8}]
    cfg = dict(flip_test=True, post_process="default", shift_heatmap=False, target_type="GaussianHeatmap", modulate_kernel=11, use_udp=True)
    res = decode_topdown(metas, maps, cfg)
    c32 = np.array([[96.5, 128.0], [50.0, 60.0]], np.float32); s32 = np.array([[192.0, 256.0], [120.0, 160.0]], np.float32)
    op, om, _ = M.keypoints_from_heatmaps(maps, c32, s32, post_process="default", use_udp=True)
    assert np.array_equal(res["preds"][..., 2:3], om, equal_nan=True) and np.nanmax(np.abs(res["preds"][..., :2] - op)) < 2e-2
    assert res["image_paths"] == ["a.jpg", "b.jpg"] and res["bbox_ids"] == [7, 8]
    assert np.allclose(res["boxes"], [[96.5, 128, 192, 256, 192 * 200 * 256 * 200, 0.9], [50, 60, 120, 160, 120 * 200 * 160 * 200, 1.0]])


def test_head_alone_and_flip_test(golden_dir):
    """model.backbone / model.keypoint_head as on the reference module (vit_models/model.py:14-24), flip_back on the GPU
    (bit-exact data movement) and the flip-test average of the reference configs (flip_test=True)."""
    from easy_vitpose_b200 import ViTPose, model_cfg
    g = np.load(os.path.join(golden_dir, "fwd_s_coco.npz"))
    D, depth, heads, K, B, wseed, xseed = (int(v) for v in g["meta"])
    m = ViTPose(model_cfg("s", K), max_batch=4)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in O.make_state_dict(D, depth, K, wseed, peaky=float(g["peaky"]), bumps=True).items()})
    m.to("cuda:0")
    x = torch.from_numpy(O.make_crops(3, 31)).cuda()
    hm = m(x)
    feats = m.backbone(x)
    assert tuple(feats.shape) == (3, D, 16, 12)
    assert torch.equal(m.keypoint_head(feats), hm)                                     # backbone -> head == forward, bit for bit
    assert np.array_equal(m.keypoint_head.inference_model(feats), hm.cpu().numpy())
    hn = hm.cpu().numpy()
    for shift in (False, True):
        m.keypoint_head.test_cfg = {"shift_heatmap": shift}
        assert np.array_equal(m.keypoint_head.inference_model(feats, M.COCO_FLIP_PAIRS), M.flip_back(hn, M.COCO_FLIP_PAIRS, shift))
    ft = m.forward_flip_test(x, M.COCO_FLIP_PAIRS)
    hf = m(torch.flip(x, dims=[3])).cpu().numpy()
    assert np.array_equal(ft.cpu().numpy(), (hn + M.flip_back(hf, M.COCO_FLIP_PAIRS)) * np.float32(0.5))
    with pytest.raises(ValueError):
        m.keypoint_head(feats[:, :-1])
