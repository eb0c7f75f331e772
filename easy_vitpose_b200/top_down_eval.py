"""keypoints_from_heatmaps on the GPU (the branch VitInference takes), same call shape as the
reference function at easy_ViTPose/vit_utils/top_down_eval.py:493-641.

Only `unbiased=True, use_udp=True, target_type='GaussianHeatmap', kernel=11` is built -- the branch
at :586-589 that easy_ViTPose/inference.py:200-203 selects; every other combination raises instead of
silently computing something else.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

__all__ = ["keypoints_from_heatmaps", "decode_heatmaps"]


def decode_heatmaps(heatmaps: torch.Tensor, org_wh: torch.Tensor, wrap_batch: bool = False):
    """heatmaps f32 CUDA [N,K,64,48], org_wh i32 [N,2] (crop width,height) -> (kpts [N,K,3] rows (y,x,score), idx [N,K])."""
    if not heatmaps.is_cuda:
        raise RuntimeError("decode_heatmaps needs a CUDA tensor: there is no CPU path")
    if heatmaps.dim() != 4 or tuple(heatmaps.shape[2:]) != (64, 48):
        raise ValueError(f"expected [N,K,64,48], got {tuple(heatmaps.shape)}")
    hm = heatmaps.to(torch.float32).contiguous()
    N, K = hm.shape[:2]
    org = torch.as_tensor(org_wh).to(device=hm.device, dtype=torch.int32).contiguous()
    if tuple(org.shape) != (N, 2):
        raise ValueError(f"org_wh must be [N,2], got {tuple(org.shape)}")
    kp = torch.empty((N, K, 3), dtype=torch.float32, device=hm.device)
    idx = torch.empty((N, K), dtype=torch.int32, device=hm.device)
    with torch.cuda.device(hm.device):
        st = C.c_void_p(torch.cuda.current_stream(hm.device).cuda_stream)
        _lib.check(_lib.lib().vpb_decode(C.c_void_p(hm.data_ptr()), N, K, C.c_void_p(org.data_ptr()), C.c_void_p(kp.data_ptr()),
                                         C.c_void_p(idx.data_ptr()), 1 if wrap_batch else 0, st))
    return kp, idx


def keypoints_from_heatmaps(heatmaps, center, scale, unbiased=False, post_process="default", kernel=11,
                            valid_radius_factor=0.0546875, use_udp=False, target_type="GaussianHeatmap"):
    """Reference signature; returns (preds [N,K,2] (x,y) float32, maxvals [N,K,1] float32) as numpy arrays.

    `center` must be scale//2-style integers and `scale` the crop (w,h) exactly as VitInference.postprocess
    passes them (inference.py:200-203): the kernel derives the centre as scale // 2.
    """
    if not (unbiased and use_udp) or str(target_type).lower() != "gaussianheatmap" or kernel != 11 \
            or post_process not in ("default", "unbiased", True):
        raise NotImplementedError("only unbiased=True, use_udp=True, GaussianHeatmap, kernel=11 is implemented on the GPU path")
    hm = heatmaps if isinstance(heatmaps, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(heatmaps, np.float32))
    if not hm.is_cuda:
        hm = hm.cuda()
    scale = np.asarray(scale)
    center = np.asarray(center)
    if scale.shape != (hm.shape[0], 2) or center.shape != scale.shape:
        raise ValueError("center and scale must be [N,2]")
    if not np.array_equal(center, scale // 2):
        raise NotImplementedError("center must equal scale // 2 (what VitInference.postprocess passes)")
    kp, _ = decode_heatmaps(hm, torch.from_numpy(scale.astype(np.int32)), wrap_batch=True)
    kp = kp.cpu().numpy()
    return np.ascontiguousarray(kp[:, :, 1::-1]), kp[:, :, 2:3].copy()
