"""CPU: the host logic of `frame_inference` (the batched replacement of VitInference.inference)
with a fake engine: detector cadence, confidence gate, tracker ids, box rounding, save_state fields.  The reference lines it
mirrors: easy_ViTPose/inference.py:232-281."""
import types

import numpy as np

from easy_vitpose_b200.inference import frame_inference


class _FakeEngine:
    num_keypoints = 2

    def __init__(self):
        self.calls = []

    def infer_frame_host(self, img, bboxes):
        bb = np.asarray(bboxes)
        self.calls.append(bb.copy())
        kp = np.zeros((len(bb), 2, 3), np.float32)
        kp[:, :, 0] = bb[:, 1:2]          # encode the box so that the dict order can be checked
        kp[:, :, 1] = bb[:, 0:1]
        return kp, np.zeros((len(bb), 2), np.int32)


def _vi(rows, tracker=None, yolo_step=1, save_state=True):
    seen = []

    def yolo(img, **kw):
        seen.append(kw)
        data = types.SimpleNamespace(cpu=lambda: types.SimpleNamespace(numpy=lambda: np.asarray(rows, np.float32)))
        return [types.SimpleNamespace(boxes=types.SimpleNamespace(data=data))]

    engine = _FakeEngine()
    vi = types.SimpleNamespace(tracker=tracker, frame_counter=0, yolo_step=yolo_step, yolo=yolo, yolo_size=320, device="cuda:1",
                               yolo_classes=[0], save_state=save_state, _b200=types.SimpleNamespace(model=engine))
    return vi, engine, seen


def test_detector_gate_rounding_and_state():
    rows = [[10.5, 20.5, 60.4, 90.6, 0.9, 0], [5, 5, 50, 50, 0.35, 0], [100.5, 11.5, 140.5, 70.5, 0.36, 0]]
    vi, engine, seen = _vi(rows)
    frame = np.zeros((120, 160, 3), np.uint8)
    out = frame_inference(vi, frame)
    assert vi.frame_counter == 1 and len(seen) == 1 and seen[0]["device"] == "cuda:1" and seen[0]["classes"] == [0]
    # the 0.35 row is dropped (strict >), .5 coordinates round half to even like numpy (inference.py:240,253)
    assert np.array_equal(engine.calls[0], np.array([[10, 20, 60, 91], [100, 12, 140, 70]]))
    assert sorted(out) == [0, 1] and out[0][0, 0] == 20 and out[1][0, 1] == 100
    boxes, ids, scores = vi._tracker_res
    assert np.array_equal(boxes, np.array([[0, 10, 70, 101], [90, 2, 150, 80]]))        # padded by 10 and clipped, as the reference leaves them
    assert ids == [0, 1] and np.allclose(scores, [0.9, 0.36]) and vi._img is frame
    assert set(vi._keypoints) == {0, 1} and np.allclose([vi._scores_bbox[0], vi._scores_bbox[1]], [0.9, 0.36])


def test_no_detections_and_no_state():
    vi, engine, _ = _vi(np.zeros((0, 6)), save_state=False)
    out = frame_inference(vi, np.zeros((50, 60, 3), np.uint8))
    assert out == {} and engine.calls[0].shape == (0, 4) and not hasattr(vi, "_img")


def test_tracker_cadence_and_ids():
    class Tracker:                                   # SORT's contract: update(dets [n,5]) -> rows [x0,y0,x1,y1,score,id]
        def __init__(self):
            self.updates = []

        def update(self, dets):
            self.updates.append(len(dets))
            return np.array([[1.2, 2.2, 30.7, 40.7, 0.8, 7], [50, 5, 90, 45, 0.7, 3]])

    trk = Tracker()
    vi, engine, seen = _vi([[0, 0, 10, 10, 0.9, 0]], tracker=trk, yolo_step=4)
    frame = np.zeros((64, 96, 3), np.uint8)
    outs = [frame_inference(vi, frame) for _ in range(9)]
    # detector on frames 0,1,2 (counter < 3), then every yolo_step-th frame: 4, 8 (inference.py:234-236)
    assert len(seen) == 5 and trk.updates == [1, 1, 1, 0, 1, 0, 0, 0, 1]
    assert all(sorted(o) == [3, 7] for o in outs)
    assert np.array_equal(engine.calls[0], np.array([[1, 2, 31, 41], [50, 5, 90, 45]]))
    assert vi._tracker_res[1] == [7, 3]
