"""easy_vitpose_b200: the ViTPose crop path of JunkyByte/easy_ViTPose on B200 (sm_100a).

    crops [B,3,256,192] -> ViT -> TopdownHeatmapSimpleHead -> heatmaps [B,K,64,48] -> keypoints [B,K,3]

csrc/ holds the hand-written CUDA (tcgen05 GEMM + attention, LayerNorm, gathers, decode) and the C ABI
(include/vitpose_b200.h); the Python modules mirror the reference's interface for this path:
model.ViTPose, top_down_eval.keypoints_from_heatmaps, inference.install / B200PoseBackend.
"""
from . import distributed  # noqa: F401
from .configs import data_cfg, dyn_model_import, model_cfg  # noqa: F401
from .inference import B200PoseBackend, install  # noqa: F401
from .model import ViTPose  # noqa: F401
from .top_down_eval import decode_heatmaps, decode_topdown, keypoints_from_heatmaps  # noqa: F401
