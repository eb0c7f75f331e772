"""COCO keypoint evaluation (OKS matching, AP / AR)  --  TEST INFRASTRUCTURE ONLY.

The reference's accuracy harness, evaluation_on_coco.py:69-81, hands its results to `pycocotools.cocoeval.COCOeval`
(iouType 'keypoints').  pycocotools (pinned by the reference at 2.0.8, requirements.txt:32) is a third-party
dependency that is ABSENT from this image and from /root/reference, and there is no network: this file restates its
published algorithm (cocoapi PythonAPI/pycocotools/cocoeval.py: _prepare, computeOks, evaluateImg, accumulate,
summarize, Params.setKpParams; coco.py: loadRes for keypoint results) in numpy.

Parity status: PARITY UNPINNED against pycocotools itself (it cannot be run here).  What pins it instead:
known-answer cases in tests/test_coco_oks_eval.py (perfect predictions -> AP 1; hand-computed OKS values; a missed
person; score ordering; area ranges).  It is used on BOTH sides of the f3 comparison (reference results vs engine
results), so an evaluator deviation would shift both APs alike; the test bar is their difference.
"""
from __future__ import annotations

import numpy as np

# Params.setKpParams: cocoeval.py
KPT_OKS_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89]) / 10.0
IOU_THRS = np.linspace(.5, 0.95, int(np.round((0.95 - .5) / .05)) + 1, endpoint=True)
REC_THRS = np.linspace(.0, 1.00, int(np.round((1.00 - .0) / .01)) + 1, endpoint=True)
MAX_DETS = 20
AREA_RNG = {"all": (0 ** 2, 1e5 ** 2), "medium": (32 ** 2, 96 ** 2), "large": (96 ** 2, 1e5 ** 2)}


def load_results(results: list[dict]) -> list[dict]:
    """COCO.loadRes for keypoint results: area and bbox come from the keypoint extent; ids are 1-based in list order."""
    out = []
    for i, r in enumerate(results):
        d = dict(r)
        s = d["keypoints"]
        x, y = s[0::3], s[1::3]
        x0, x1, y0, y1 = np.min(x), np.max(x), np.min(y), np.max(y)
        d["area"] = float((x1 - x0) * (y1 - y0))
        d["id"] = i + 1
        d["bbox"] = [x0, y0, x1 - x0, y1 - y0]
        out.append(d)
    return out


def compute_oks(gts: list[dict], dts: list[dict], sigmas: np.ndarray = KPT_OKS_SIGMAS) -> np.ndarray:
    """COCOeval.computeOks for one image: [len(dts), len(gts)], dts already sorted by descending score and truncated."""
    if len(gts) == 0 or len(dts) == 0:
        return np.zeros((len(dts), len(gts)))
    ious = np.zeros((len(dts), len(gts)))
    vars_ = (sigmas * 2) ** 2
    k = len(sigmas)
    for j, gt in enumerate(gts):
        g = np.array(gt["keypoints"], dtype=np.float64)
        xg, yg, vg = g[0::3], g[1::3], g[2::3]
        k1 = np.count_nonzero(vg > 0)
        bb = gt["bbox"]
        x0, x1 = bb[0] - bb[2], bb[0] + bb[2] * 2
        y0, y1 = bb[1] - bb[3], bb[1] + bb[3] * 2
        for i, dt in enumerate(dts):
            d = np.array(dt["keypoints"], dtype=np.float64)
            xd, yd = d[0::3], d[1::3]
            if k1 > 0:
                dx, dy = xd - xg, yd - yg
            else:
                z = np.zeros(k)
                dx = np.max((z, x0 - xd), axis=0) + np.max((z, xd - x1), axis=0)
                dy = np.max((z, y0 - yd), axis=0) + np.max((z, yd - y1), axis=0)
            e = (dx ** 2 + dy ** 2) / vars_ / (gt["area"] + np.spacing(1)) / 2
            if k1 > 0:
                e = e[vg > 0]
            ious[i, j] = np.sum(np.exp(-e)) / e.shape[0]
    return ious


def _evaluate_img(gts: list[dict], dts: list[dict], a_rng, sigmas) -> dict | None:
    """COCOeval.evaluateImg for one image and one area range (one category, maxDet = 20)."""
    if len(gts) == 0 and len(dts) == 0:
        return None
    for g in gts:
        ign = bool(g.get("ignore", 0)) or bool(g.get("iscrowd", 0))
        ign = ign or g.get("num_keypoints", 1) == 0                 # _prepare, keypoints
        g["_ignore"] = 1 if (ign or g["area"] < a_rng[0] or g["area"] > a_rng[1]) else 0
    dtind = np.argsort([-d["score"] for d in dts], kind="mergesort")[:MAX_DETS]
    dts_all = [dts[i] for i in dtind]
    ious_all = compute_oks(gts, dts_all, sigmas)                     # computed on the unsorted gts, then re-indexed
    gtind = np.argsort([g["_ignore"] for g in gts], kind="mergesort")
    gt = [gts[i] for i in gtind]
    dt = dts_all
    iscrowd = [int(g.get("iscrowd", 0)) for g in gt]
    ious = ious_all[:, gtind] if len(ious_all) > 0 else ious_all
    T, G, D = len(IOU_THRS), len(gt), len(dt)
    gtm, dtm = np.zeros((T, G)), np.zeros((T, D))
    gt_ig = np.array([g["_ignore"] for g in gt])
    dt_ig = np.zeros((T, D))
    if len(ious) != 0:
        for tind, t in enumerate(IOU_THRS):
            for dind in range(D):
                iou = min([t, 1 - 1e-10])
                m = -1
                for gind in range(G):
                    if gtm[tind, gind] > 0 and not iscrowd[gind]:
                        continue
                    if m > -1 and gt_ig[m] == 0 and gt_ig[gind] == 1:
                        break
                    if ious[dind, gind] < iou:
                        continue
                    iou = ious[dind, gind]
                    m = gind
                if m == -1:
                    continue
                dt_ig[tind, dind] = gt_ig[m]
                dtm[tind, dind] = gt[m]["id"]
                gtm[tind, m] = dt[dind]["id"]
    a = np.array([d["area"] < a_rng[0] or d["area"] > a_rng[1] for d in dt]).reshape((1, len(dt)))
    dt_ig = np.logical_or(dt_ig, np.logical_and(dtm == 0, np.repeat(a, T, 0)))
    return {"dtMatches": dtm, "dtScores": [d["score"] for d in dt], "gtIgnore": gt_ig, "dtIgnore": dt_ig}


def evaluate(gt_annotations: list[dict], results: list[dict], image_ids, sigmas: np.ndarray = KPT_OKS_SIGMAS) -> dict:
    """COCOeval(cocoGt, cocoGt.loadRes(results), 'keypoints') .evaluate() .accumulate() .summarize() for category 1:
    returns the ten summary numbers under their usual names."""
    dts_all = load_results(results)
    stats = {}
    prec = {}
    rec = {}
    for a_name, a_rng in AREA_RNG.items():
        evals = []
        for img in image_ids:
            gts = [dict(g) for g in gt_annotations if g["image_id"] == img and g.get("category_id", 1) == 1]
            dts = [d for d in dts_all if d["image_id"] == img and d.get("category_id", 1) == 1]
            e = _evaluate_img(gts, dts, a_rng, sigmas)
            if e is not None:
                evals.append(e)
        T, R = len(IOU_THRS), len(REC_THRS)
        precision = -np.ones((T, R))
        recall = -np.ones((T,))
        if evals:
            dt_scores = np.concatenate([e["dtScores"][0:MAX_DETS] for e in evals])
            inds = np.argsort(-dt_scores, kind="mergesort")
            dtm = np.concatenate([e["dtMatches"][:, 0:MAX_DETS] for e in evals], axis=1)[:, inds]
            dt_ig = np.concatenate([e["dtIgnore"][:, 0:MAX_DETS] for e in evals], axis=1)[:, inds]
            gt_ig = np.concatenate([e["gtIgnore"] for e in evals])
            npig = np.count_nonzero(gt_ig == 0)
            if npig > 0:
                tps = np.logical_and(dtm, np.logical_not(dt_ig))
                fps = np.logical_and(np.logical_not(dtm), np.logical_not(dt_ig))
                tp_sum = np.cumsum(tps, axis=1).astype(dtype=float)
                fp_sum = np.cumsum(fps, axis=1).astype(dtype=float)
                for t, (tp, fp) in enumerate(zip(tp_sum, fp_sum)):
                    nd = len(tp)
                    rc = tp / npig
                    pr = tp / (fp + tp + np.spacing(1))
                    q = np.zeros((R,))
                    recall[t] = rc[-1] if nd else 0
                    pr = pr.tolist()
                    for i in range(nd - 1, 0, -1):
                        if pr[i] > pr[i - 1]:
                            pr[i - 1] = pr[i]
                    idx = np.searchsorted(rc, REC_THRS, side="left")
                    for ri, pi in enumerate(idx):
                        if pi < nd:
                            q[ri] = pr[pi]
                    precision[t] = q
        prec[a_name], rec[a_name] = precision, recall

    def _mean(arr):
        arr = arr[arr > -1]
        return float(np.mean(arr)) if arr.size else -1.0
    t50, t75 = int(np.where(np.isclose(IOU_THRS, 0.5))[0][0]), int(np.where(np.isclose(IOU_THRS, 0.75))[0][0])
    stats["AP"] = _mean(prec["all"])
    stats["AP50"] = _mean(prec["all"][t50])
    stats["AP75"] = _mean(prec["all"][t75])
    stats["AP_medium"] = _mean(prec["medium"])
    stats["AP_large"] = _mean(prec["large"])
    stats["AR"] = _mean(rec["all"])
    stats["AR50"] = _mean(rec["all"][t50:t50 + 1])
    stats["AR75"] = _mean(rec["all"][t75:t75 + 1])
    stats["AR_medium"] = _mean(rec["medium"])
    stats["AR_large"] = _mean(rec["large"])
    return stats


def results_from_frame_keypoints(image_id: int, frame_keypoints: dict, scores_bbox: dict) -> list[dict]:
    """The result records evaluation_on_coco.py:52-66 builds from `model.inference(img)` and `model._scores_bbox`:
    keypoints as rounded (x, y, 0) triples, score = detector confidence."""
    out = []
    for key in frame_keypoints:
        kps = []
        for k in frame_keypoints[key]:
            kps.append(float(round(float(k[1]), 0)))
            kps.append(float(round(float(k[0]), 0)))
            kps.append(0)
        out.append({"image_id": image_id, "category_id": 1, "score": scores_bbox[key], "bbox": [], "keypoints": kps})
    return out
