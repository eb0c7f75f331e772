"""Import the UNMODIFIED reference modules from /root/reference (or baseline/_ref)  --  TEST INFRASTRUCTURE ONLY.

Used by oracle/make_golden.py (and by tests that are skipped when the reference tree is
absent, i.e. on the GPU box) to pin oracle/vitpose_oracle.py against the reference's own
arithmetic.  Nothing on the product path imports this.

The reference cannot be imported as-is in this image: vit_utils/__init__.py:4 pulls in
vit_utils/visualization.py, which imports matplotlib.pyplot and ffmpeg (both absent).
Empty module stubs for those names are enough (probed).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    """Where the UNMODIFIED reference package lives: $EASY_VITPOSE_REF, the read-only mount of the build container, or the
    copy `pip install --target baseline/_ref /root/reference` leaves in the repo (git-ignored; it travels to the GPU box,
    which has no /root/reference)."""
    cands = [os.environ.get("EASY_VITPOSE_REF"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "easy_ViTPose", "vit_models")):
            return c
    return cands[1]


REF_ROOT = _find_root()
REF_PKG = os.path.join(REF_ROOT, "easy_ViTPose")


def available() -> bool:
    return os.path.isdir(REF_PKG)


def load():
    """Returns a namespace with ViTPose, dyn_model_import, keypoints_from_heatmaps,
    transform_preds from the reference tree."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True            # the reference mount is read-only
    for name in ("matplotlib", "matplotlib.pyplot", "ffmpeg"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if REF_PKG not in sys.path:
        sys.path.insert(0, REF_PKG)
    ns = types.SimpleNamespace()
    ns.ViTPose = importlib.import_module("vit_models.model").ViTPose
    ns.dyn_model_import = importlib.import_module("vit_utils.util").dyn_model_import
    tde = importlib.import_module("vit_utils.top_down_eval")
    ns.keypoints_from_heatmaps = tde.keypoints_from_heatmaps
    ns.transform_preds = importlib.import_module("vit_utils.post_processing.post_transforms").transform_preds
    return ns


def postprocess(ns, heatmaps, org_w, org_h):
    """What VitInference.postprocess does (easy_ViTPose/inference.py:187-205), calling the
    reference's keypoints_from_heatmaps.  VitInference itself is not importable here
    (top-level `from ultralytics import YOLO`, inference.py:10)."""
    import warnings

    import numpy as np
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        points, prob = ns.keypoints_from_heatmaps(
            heatmaps=heatmaps, center=np.array([[org_w // 2, org_h // 2]]),
            scale=np.array([[org_w, org_h]]), unbiased=True, use_udp=True)
    return np.concatenate([points[:, :, ::-1], prob], axis=2)


class _Permissive(types.ModuleType):
    """Stand-in for an absent third-party module: any attribute is `object` (enough for `from x import Y` at import time)."""
    __path__: list = []

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return object


def load_vitinference():
    """The UNMODIFIED `easy_ViTPose.inference` module (VitInference, pad_image, MEAN, STD).  Its import chain needs
    ultralytics, matplotlib, skimage, filterpy, ffmpeg (inference.py:10, sort.py:22-29, visualization.py) -- none is
    touched by the per-person pose loop, so permissive stubs are enough.  The detector is supplied by the caller
    (oracle/make_golden_frames.py uses a stub that returns fixed boxes)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "ffmpeg", "ultralytics", "skimage", "skimage.io",
                 "filterpy", "filterpy.kalman", "lap"):
        if name not in sys.modules or not hasattr(sys.modules[name], "__path__") and name in ("matplotlib", "skimage", "filterpy"):
            try:
                if isinstance(sys.modules.get(name), types.ModuleType) and not getattr(sys.modules[name], "__file__", None):
                    raise ImportError(name)                      # an empty stub left by load(): replace it
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Permissive(name)
    if hasattr(sys.modules["matplotlib"], "__path__") and not getattr(sys.modules["matplotlib"], "__file__", None):
        sys.modules["matplotlib"].use = lambda *a, **k: None
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return importlib.import_module("easy_ViTPose.inference")
