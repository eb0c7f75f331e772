"""CPU: known-answer tests for oracle/coco_oks_eval.py (the restatement of pycocotools' keypoint COCOeval that the f3
accuracy harness uses on both sides; pycocotools itself is absent from the image)."""
import numpy as np

from oracle import coco_oks_eval as E


def _person(img, pid, cx, cy, size, visible=17):
    rs = np.random.RandomState(pid)
    xs = cx + rs.uniform(-size / 2, size / 2, 17)
    ys = cy + rs.uniform(-size, size, 17)
    v = np.array([2] * visible + [0] * (17 - visible))
    kp = np.stack([xs, ys, v], 1).reshape(-1).tolist()
    return {"id": pid, "image_id": img, "category_id": 1, "iscrowd": 0, "num_keypoints": int(visible), "keypoints": kp,
            "bbox": [cx - size / 2, cy - size, size, 2 * size], "area": float(size * 2 * size)}


def _result(gt, score, shift=(0.0, 0.0)):
    kp = np.array(gt["keypoints"]).reshape(17, 3).copy()
    kp[:, 0] += shift[0]; kp[:, 1] += shift[1]; kp[:, 2] = 0
    return {"image_id": gt["image_id"], "category_id": 1, "score": score, "bbox": [], "keypoints": kp.reshape(-1).tolist()}


def test_perfect_predictions_score_one():
    gts = [_person(1, 1, 100, 100, 60), _person(1, 2, 300, 120, 80), _person(2, 3, 200, 200, 120)]
    res = [_result(g, 0.9 - 0.1 * i) for i, g in enumerate(gts)]
    s = E.evaluate(gts, res, [1, 2])
    assert abs(s["AP"] - 1.0) < 1e-9 and abs(s["AP50"] - 1.0) < 1e-9 and s["AR"] == 1.0


def test_oks_of_a_uniform_shift_matches_the_formula():
    g = _person(1, 1, 100, 100, 60)
    d = E.load_results([_result(g, 0.9, shift=(3.0, 4.0))])
    oks = E.compute_oks([g], d)[0, 0]
    want = np.mean(np.exp(-(25.0) / ((E.KPT_OKS_SIGMAS * 2) ** 2) / (g["area"] + np.spacing(1)) / 2))
    assert abs(oks - want) < 1e-12
    # only labelled keypoints count
    g2 = _person(1, 2, 100, 100, 60, visible=5)
    oks2 = E.compute_oks([g2], E.load_results([_result(g2, 0.9, shift=(3.0, 4.0))]))[0, 0]
    want2 = np.mean(np.exp(-(25.0) / ((E.KPT_OKS_SIGMAS[:5] * 2) ** 2) / (g2["area"] + np.spacing(1)) / 2))
    assert abs(oks2 - want2) < 1e-12


def test_threshold_sweep_and_a_missed_person():
    """Two people, one detected with OKS ~0.72, one missed: AP50 = 0.5-recall plateau, thresholds above the OKS give 0."""
    g1, g2 = _person(1, 1, 100, 100, 60), _person(1, 2, 300, 120, 80)
    sh = 1.0
    while E.compute_oks([g1], E.load_results([_result(g1, 0.9, shift=(sh, 0.0))]))[0, 0] > 0.72:
        sh += 0.25
    oks = E.compute_oks([g1], E.load_results([_result(g1, 0.9, shift=(sh, 0.0))]))[0, 0]
    s = E.evaluate([g1, g2], [_result(g1, 0.9, shift=(sh, 0.0))], [1])
    n_pass = int((E.IOU_THRS <= oks + 1e-12).sum())                      # thresholds the detection passes
    # at a passing threshold: precision 1 up to recall 0.5 (51 of 101 recall points), 0 beyond; otherwise 0 everywhere
    assert abs(s["AP"] - n_pass * (51 / 101) / 10) < 1e-9
    assert abs(s["AP50"] - 51 / 101) < 1e-9
    assert abs(s["AR"] - n_pass * 0.5 / 10) < 1e-9


def test_score_order_decides_precision():
    """A false positive that outranks the true positive halves the precision at every recall point."""
    g = _person(1, 1, 100, 100, 60)
    far = _result(g, 0.95, shift=(500.0, 500.0))
    good = _result(g, 0.5)
    s_hi = E.evaluate([g], [far, good], [1])
    s_lo = E.evaluate([g], [dict(far, score=0.1), good], [1])
    assert abs(s_hi["AP"] - 0.5) < 1e-9 and abs(s_lo["AP"] - 1.0) < 1e-9


def test_area_ranges():
    small, large = _person(1, 1, 100, 100, 30), _person(1, 2, 300, 200, 100)      # areas 1800 (medium) and 20000 (large)
    res = [_result(small, 0.9), _result(large, 0.8, shift=(400.0, 0.0))]
    s = E.evaluate([small, large], res, [1])
    assert abs(s["AP_medium"] - 1.0) < 1e-9 and s["AP_large"] == 0.0


def test_result_records_follow_the_reference_script():
    kp = {7: np.array([[10.4, 20.6, 0.9], [11.5, 22.5, 0.8]], np.float32)}
    r = E.results_from_frame_keypoints(42, kp, {7: 0.77})
    assert r == [{"image_id": 42, "category_id": 1, "score": 0.77, "bbox": [], "keypoints": [21.0, 10.0, 0, 22.0, 12.0, 0]}]
