"""ViTPose(cfg): the reference's model object (easy_ViTPose/vit_models/model.py:10-24) backed by the
sm_100a engine.  Same constructor argument, same state_dict key contract, same forward signature;
the arithmetic runs in libvitpose_b200.so (bf16 tensor-core GEMMs, fp32 residual stream/softmax/LN).

torch is used for what it is good at here: owning device memory and the current stream.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np
import torch

from . import _lib

__all__ = ["ViTPose"]

IMG_H, IMG_W, HM_H, HM_W = 256, 192, 64, 48


def _expected_shapes(D: int, depth: int, K: int) -> "OrderedDict[str, tuple]":
    """state_dict contract of the reference ViTPose."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["backbone.pos_embed"] = (1, 193, D)
    s["backbone.patch_embed.proj.weight"] = (D, 3, 16, 16)
    s["backbone.patch_embed.proj.bias"] = (D,)
    for i in range(depth):
        p = f"backbone.blocks.{i}."
        s[p + "norm1.weight"] = (D,); s[p + "norm1.bias"] = (D,)
        s[p + "attn.qkv.weight"] = (3 * D, D); s[p + "attn.qkv.bias"] = (3 * D,)
        s[p + "attn.proj.weight"] = (D, D); s[p + "attn.proj.bias"] = (D,)
        s[p + "norm2.weight"] = (D,); s[p + "norm2.bias"] = (D,)
        s[p + "mlp.fc1.weight"] = (4 * D, D); s[p + "mlp.fc1.bias"] = (4 * D,)
        s[p + "mlp.fc2.weight"] = (D, 4 * D); s[p + "mlp.fc2.bias"] = (D,)
    s["backbone.last_norm.weight"] = (D,); s["backbone.last_norm.bias"] = (D,)
    cin = D
    for li in (0, 3):
        s[f"keypoint_head.deconv_layers.{li}.weight"] = (cin, 256, 4, 4)
        for n in ("weight", "bias", "running_mean", "running_var"):
            s[f"keypoint_head.deconv_layers.{li + 1}.{n}"] = (256,)
        s[f"keypoint_head.deconv_layers.{li + 1}.num_batches_tracked"] = ()
        cin = 256
    s["keypoint_head.final_layer.weight"] = (K, 256, 1, 1)
    s["keypoint_head.final_layer.bias"] = (K,)
    return s


class _Backbone:
    """`model.backbone` of the reference ViTPose (vit_models/model.py:14): callable, [B,3,256,192] -> [B,D,16,12]."""

    def __init__(self, owner):
        self._owner = owner
        self.num_heads = owner.num_heads if hasattr(owner, "num_heads") else None

    def __call__(self, x):
        return self._owner.forward_features(x)

    forward = __call__


class _Head:
    """`model.keypoint_head` (vit_models/model.py:15, head/topdown_heatmap_simple_head.py): forward and inference_model."""

    def __init__(self, owner):
        self._owner = owner
        self.target_type = "GaussianHeatmap"
        self.test_cfg = {}

    def __call__(self, features):
        return self._owner.head_forward(features)

    forward = __call__

    def inference_model(self, x, flip_pairs=None):
        """head/topdown_heatmap_simple_head.py:195-218: numpy heatmaps, flipped back (and shifted by one pixel when
        test_cfg['shift_heatmap']) if flip_pairs is given."""
        out = self._owner.head_forward(x)
        if flip_pairs is not None:
            out = self._owner.flip_back(out, flip_pairs, bool(self.test_cfg.get("shift_heatmap", False)))
        return out.cpu().numpy()


class ViTPose:
    """Drop-in for the reference `ViTPose(cfg)` on the inference path.

    cfg is the reference's model dict (cfg['backbone'], cfg['keypoint_head']; 'type' keys ignored,
    model.py:14-15).  Only the configurations the reference ships are accepted: patch 16, 256x192,
    mlp_ratio 4, qkv_bias, two 4x4 deconvs of 256 filters and a 1x1 final conv.
    """

    def __init__(self, cfg: dict, max_batch: int = 64, device: "int | str | torch.device | None" = None) -> None:
        bb = {k: v for k, v in cfg["backbone"].items() if k != "type"}
        hd = {k: v for k, v in cfg["keypoint_head"].items() if k != "type"}
        if tuple(bb.get("img_size", (256, 192))) != (256, 192) or bb.get("patch_size", 16) != 16 or bb.get("ratio", 1) != 1:
            raise ValueError("only img_size=(256,192), patch_size=16, ratio=1 (the reference's configs) are built")
        if bb.get("mlp_ratio", 4) != 4 or not bb.get("qkv_bias", False):
            raise ValueError("only mlp_ratio=4, qkv_bias=True (the reference's configs) are built")
        if hd.get("num_deconv_layers", 3) != 2 or tuple(hd.get("num_deconv_filters", ())) != (256, 256) \
                or tuple(hd.get("num_deconv_kernels", ())) != (4, 4) or (hd.get("extra") or {}).get("final_conv_kernel", 1) != 1:
            raise ValueError("only the 2x deconv(256,4x4) + 1x1 conv head of the reference's configs is built")
        self.embed_dim = int(bb["embed_dim"]); self.depth = int(bb["depth"]); self.num_heads = int(bb["num_heads"])
        self.num_keypoints = int(hd["out_channels"])
        if int(hd["in_channels"]) != self.embed_dim:
            raise ValueError("keypoint_head.in_channels must equal backbone.embed_dim")
        self.max_batch = int(max_batch)
        self.training = False
        self._cfg = cfg
        self._handle = C.c_void_p()
        self._loaded = False
        self._state: "OrderedDict[str, torch.Tensor] | None" = None
        self._device = None
        self._side = None
        self.backbone = _Backbone(self)              # model.backbone(x) / model.keypoint_head(f), as on the reference module
        self.keypoint_head = _Head(self)
        if device is not None:
            self.to(device)

    # ---------------------------------------------------------------- nn.Module-like surface
    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise RuntimeError("the B200 engine is inference-only (eval mode); training stays with the reference")
        return self

    def to(self, device):
        dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        if dev.type != "cuda":
            raise RuntimeError(f"ViTPose(B200) has no {dev.type} path: it needs a CUDA sm_100 device")
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        if self._device is not None and self._device != index and self._handle:
            raise RuntimeError("engine already lives on another device")
        self._device = index
        if self._state is not None and not self._loaded:
            self._upload()
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device if device is not None else torch.cuda.current_device()))

    def state_dict(self):
        if self._state is None:
            raise RuntimeError("no weights loaded")
        return OrderedDict((k, v.clone()) for k, v in self._state.items())

    def load_state_dict(self, state_dict, strict: bool = True):
        """Same contract as nn.Module.load_state_dict(strict=True): missing / unexpected keys and
        shape mismatches raise (easy_ViTPose/inference.py:162-166 calls it exactly like this)."""
        if "state_dict" in state_dict and not any(k.startswith("backbone.") for k in state_dict):
            state_dict = state_dict["state_dict"]
        exp = _expected_shapes(self.embed_dim, self.depth, self.num_keypoints)
        missing = [k for k in exp if k not in state_dict and not k.endswith("num_batches_tracked")]
        unexpected = [k for k in state_dict if k not in exp]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for ViTPose: missing keys {missing[:5]}"
                               f"{'...' if len(missing) > 5 else ''}, unexpected keys {unexpected[:5]}")
        st: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        for k, shape in exp.items():
            if k not in state_dict:
                if k.endswith("num_batches_tracked"):
                    st[k] = torch.zeros((), dtype=torch.int64)
                    continue
                raise RuntimeError(f"missing key {k}")
            v = state_dict[k]
            v = torch.as_tensor(np.asarray(v)) if not isinstance(v, torch.Tensor) else v
            if tuple(v.shape) != tuple(shape):
                raise RuntimeError(f"size mismatch for {k}: copying a param with shape {tuple(v.shape)}, expected {tuple(shape)}")
            st[k] = v.detach().to("cpu").contiguous()
        if self._loaded:
            raise RuntimeError("weights already packed on the device; create a new ViTPose to load another checkpoint")
        self._state = st
        if self._device is not None:
            self._upload()
        return self

    # ---------------------------------------------------------------- engine plumbing
    def _ensure(self):
        if not self._loaded:
            if self._state is None:
                raise RuntimeError("load_state_dict() before forward()")
            if self._device is None:
                self.to("cuda")
            else:
                self._upload()

    def _upload(self):
        L = _lib.lib()
        cfg = _lib.VpbConfig(self.embed_dim, self.depth, self.num_heads, self.num_keypoints, self.max_batch, self._device)
        with torch.cuda.device(self._device):
            _lib.check(L.vpb_create(C.byref(cfg), C.byref(self._handle)))
            for k, v in self._state.items():
                if k.endswith("num_batches_tracked"):
                    continue
                a = v.to(torch.float32).contiguous().numpy()
                _lib.check(L.vpb_load_tensor(self._handle, k.encode(), a.ctypes.data_as(C.c_void_p), a.size))
            _lib.check(L.vpb_finalize(self._handle))
        self._loaded = True

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self._device).cuda_stream)

    def _check_input(self, x: torch.Tensor) -> torch.Tensor:
        if not isinstance(x, torch.Tensor):
            raise TypeError("expected a torch.Tensor [B,3,256,192]")
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, IMG_H, IMG_W):
            raise ValueError(f"expected [B,3,256,192], got {tuple(x.shape)}")
        if x.shape[0] < 1 or x.shape[0] > self.max_batch:
            raise ValueError(f"batch {x.shape[0]} outside 1..max_batch={self.max_batch}")
        self._ensure()
        if not x.is_cuda:
            x = x.to(torch.device("cuda", self._device), non_blocking=True)
        if x.device.index != self._device:
            raise ValueError("input lives on another GPU than the engine")
        return x.to(torch.float32).contiguous()

    # ---------------------------------------------------------------- forward paths
    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[B,3,256,192] float32 -> heatmaps [B,K,64,48] float32 (model.py:23-24)."""
        x = self._check_input(x)
        out = torch.empty((x.shape[0], self.num_keypoints, HM_H, HM_W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().vpb_forward(self._handle, C.c_void_p(x.data_ptr()), x.shape[0], C.c_void_p(out.data_ptr()), self._stream()))
        return out

    __call__ = forward

    @torch.no_grad()
    def forward_features(self, x: torch.Tensor) -> torch.Tensor:
        """[B,3,256,192] -> backbone features [B,D,16,12] (model.py:20-21, vit.py:375-389)."""
        x = self._check_input(x)
        out = torch.empty((x.shape[0], self.embed_dim, 16, 12), dtype=torch.float32, device=x.device)
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().vpb_forward_features(self._handle, C.c_void_p(x.data_ptr()), x.shape[0], C.c_void_p(out.data_ptr()), self._stream()))
        return out

    @torch.no_grad()
    def head_forward(self, features: torch.Tensor) -> torch.Tensor:
        """Backbone features [B,D,16,12] -> heatmaps [B,K,64,48]: TopdownHeatmapSimpleHead.forward
        (head/topdown_heatmap_simple_head.py:188-193).  The engine's head consumes bf16 features; those returned by
        forward_features are bf16 values already, so backbone -> head in two calls equals forward()."""
        self._ensure()
        if not isinstance(features, torch.Tensor) or features.dim() != 4 or tuple(features.shape[1:]) != (self.embed_dim, 16, 12):
            raise ValueError(f"expected features [B,{self.embed_dim},16,12]")
        if features.shape[0] < 1 or features.shape[0] > self.max_batch:
            raise ValueError(f"batch {features.shape[0]} outside 1..max_batch={self.max_batch}")
        f = features.to(device=torch.device("cuda", self._device), dtype=torch.float32).contiguous()
        out = torch.empty((f.shape[0], self.num_keypoints, HM_H, HM_W), dtype=torch.float32, device=f.device)
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().vpb_head(self._handle, C.c_void_p(f.data_ptr()), f.shape[0], C.c_void_p(out.data_ptr()), self._stream()))
        return out

    @torch.no_grad()
    def flip_back(self, heatmaps: torch.Tensor, flip_pairs, shift_heatmap: bool = False) -> torch.Tensor:
        """flip_back (post_processing/post_transforms.py:110-147) + the optional shift of inference_model (:210-212) on the GPU."""
        hm = heatmaps.to(device=torch.device("cuda", self._device if self._device is not None else torch.cuda.current_device()),
                         dtype=torch.float32).contiguous()
        if hm.dim() != 4 or tuple(hm.shape[2:]) != (HM_H, HM_W):
            raise ValueError(f"expected [N,K,64,48], got {tuple(hm.shape)}")
        K = hm.shape[1]
        perm = list(range(K))
        for left, right in flip_pairs:                       # sequential, like the reference's loop over (left, right)
            perm[left], perm[right] = right, left
        pt = torch.tensor(perm, dtype=torch.int32, device=hm.device)
        out = torch.empty_like(hm)
        with torch.cuda.device(hm.device):
            _lib.check(_lib.lib().vpb_flip_back(C.c_void_p(hm.data_ptr()), hm.shape[0], K, C.c_void_p(pt.data_ptr()), 1 if shift_heatmap else 0,
                                                C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(hm.device).cuda_stream)))
        return out

    @torch.no_grad()
    def forward_flip_test(self, x: torch.Tensor, flip_pairs, shift_heatmap: bool = False) -> torch.Tensor:
        """mmpose's flip test (the `flip_test=True` of the reference configs, configs/ViTPose_common.py:124): heatmaps of the
        image and of its mirror image (flipped back, keypoint pairs swapped) averaged -- the published-AP protocol."""
        x = self._check_input(x)
        hm = self.forward(x)
        hm_f = self.flip_back(self.forward(torch.flip(x, dims=[3])), flip_pairs, shift_heatmap)
        return (hm + hm_f) * 0.5

    def _call_on_stream(self, tensors, call) -> None:
        """Runs `call(stream)` on the caller's current stream; on the legacy default stream (which cannot be captured into a
        CUDA graph) on a side stream ordered after / before it by two event waits, so small batches get graph replay."""
        with torch.cuda.device(self._device):
            cur = torch.cuda.current_stream(self._device)
            if cur.cuda_stream == 0:
                if self._side is None:
                    self._side = torch.cuda.Stream(self._device)
                self._side.wait_stream(cur)
                for t in tensors:
                    if t is not None:
                        t.record_stream(self._side)
                _lib.check(call(C.c_void_p(self._side.cuda_stream)))
                cur.wait_stream(self._side)
            else:
                _lib.check(call(C.c_void_p(cur.cuda_stream)))

    @torch.no_grad()
    def infer_crops(self, x: torch.Tensor, org_wh: torch.Tensor, return_heatmaps: bool = False):
        """Batched crops -> keypoints [B,K,3] (y, x, score) in crop pixels + flat argmax [B,K].
        org_wh int32 [B,2] = each crop's (width, height) before the resize to 192x256."""
        x = self._check_input(x)
        B = x.shape[0]
        org = torch.as_tensor(org_wh).to(device=x.device, dtype=torch.int32).contiguous()
        if tuple(org.shape) != (B, 2):
            raise ValueError(f"org_wh must be [B,2], got {tuple(org.shape)}")
        kp = torch.empty((B, self.num_keypoints, 3), dtype=torch.float32, device=x.device)
        idx = torch.empty((B, self.num_keypoints), dtype=torch.int32, device=x.device)
        hm = torch.empty((B, self.num_keypoints, HM_H, HM_W), dtype=torch.float32, device=x.device) if return_heatmaps else None
        self._call_on_stream((x, org, kp, idx, hm), lambda st: _lib.lib().vpb_infer(
            self._handle, C.c_void_p(x.data_ptr()), C.c_void_p(org.data_ptr()), B, C.c_void_p(kp.data_ptr()),
            C.c_void_p(idx.data_ptr()), C.c_void_p(hm.data_ptr()) if hm is not None else None, st))
        return (kp, idx, hm) if return_heatmaps else (kp, idx)

    def infer_host(self, crops: np.ndarray, org_wh: np.ndarray, kpts_out: np.ndarray | None = None,
                   idx_out: np.ndarray | None = None):
        """HOST buffers in, HOST keypoints out through the C ABI (vpb_infer_host): H2D + path + D2H + sync."""
        self._ensure()
        crops = np.ascontiguousarray(crops, np.float32)
        org = np.ascontiguousarray(org_wh, np.int32)
        B = crops.shape[0]
        if crops.shape[1:] != (3, IMG_H, IMG_W) or org.shape != (B, 2):
            raise ValueError("crops [B,3,256,192] float32 and org_wh [B,2] int32 expected")
        kp = kpts_out if kpts_out is not None else np.empty((B, self.num_keypoints, 3), np.float32)
        idx = idx_out if idx_out is not None else np.empty((B, self.num_keypoints), np.int32)
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().vpb_infer_host(self._handle, crops.ctypes.data_as(C.c_void_p), org.ctypes.data_as(C.c_void_p), B,
                                                 kp.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p), self._stream()))
        return kp, idx

    def submit_host(self, crops: np.ndarray, org_wh: np.ndarray, kpts_out: np.ndarray, idx_out: np.ndarray, slot: int) -> None:
        """Asynchronous vpb_submit_host: the arrays must be C-contiguous float32 / int32, stay alive and unmodified until
        wait_host(slot); pinned memory (torch.Tensor.pin_memory().numpy()) gives real copy/compute overlap."""
        self._ensure()
        B = crops.shape[0]
        if crops.dtype != np.float32 or org_wh.dtype != np.int32 or kpts_out.dtype != np.float32 or idx_out.dtype != np.int32:
            raise TypeError("submit_host takes float32 crops / keypoints and int32 org_wh / idx")
        if crops.shape[1:] != (3, IMG_H, IMG_W) or org_wh.shape != (B, 2) or kpts_out.shape != (B, self.num_keypoints, 3) \
                or idx_out.shape != (B, self.num_keypoints) or not all(a.flags.c_contiguous for a in (crops, org_wh, kpts_out, idx_out)):
            raise ValueError("submit_host: wrong shapes or non-contiguous arrays")
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().vpb_submit_host(self._handle, crops.ctypes.data_as(C.c_void_p), org_wh.ctypes.data_as(C.c_void_p), B,
                                                  kpts_out.ctypes.data_as(C.c_void_p), idx_out.ctypes.data_as(C.c_void_p), int(slot)))

    # ---------------------------------------------------------------------------------------- frame-level calls
    def _check_frame(self, frame: torch.Tensor, bboxes) -> "tuple[torch.Tensor, torch.Tensor]":
        self._ensure()
        if not isinstance(frame, torch.Tensor) or frame.dtype != torch.uint8 or frame.dim() != 3 or frame.shape[2] != 3:
            raise ValueError("frame must be a uint8 RGB tensor [H,W,3]")
        dev = torch.device("cuda", self._device)
        if not frame.is_cuda:
            frame = frame.to(dev, non_blocking=True)
        if frame.device.index != self._device:
            raise ValueError(f"frame lives on {frame.device}, the engine on cuda:{self._device}")
        bb = torch.as_tensor(bboxes)
        if bb.is_floating_point():
            bb = bb.round()                                  # easy_ViTPose/inference.py:253 (round half to even, like numpy)
        bb = bb.to(device=dev, dtype=torch.int32).reshape(-1, 4).contiguous()
        if bb.shape[0] > self.max_batch:
            raise ValueError(f"{bb.shape[0]} boxes exceed max_batch={self.max_batch}")
        return frame.contiguous(), bb

    def preprocess(self, frame: torch.Tensor, bboxes, pad_bbox: int = 10):
        """uint8 RGB frame [H,W,3] (CUDA) + boxes [n,4] (x0,y0,x1,y1) -> (crops f32 [n,3,256,192], org_wh i32 [n,2],
        offs_yx i32 [n,2]): box padding/clipping, pad_image and pre_img of the reference in one kernel
        (easy_ViTPose/inference.py:259-265,314-318).  Raises ValueError for a box that is empty after clipping."""
        frame, bb = self._check_frame(frame, bboxes)
        n = bb.shape[0]
        dev = frame.device
        crops = torch.empty((n, 3, IMG_H, IMG_W), dtype=torch.float32, device=dev)
        org = torch.empty((n, 2), dtype=torch.int32, device=dev)
        offs = torch.empty((n, 2), dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        if n:
            with torch.cuda.device(self._device):
                _lib.check(_lib.lib().vpb_preprocess(C.c_void_p(frame.data_ptr()), frame.shape[0], frame.shape[1], 0,
                                                     C.c_void_p(bb.data_ptr()), n, int(pad_bbox), C.c_void_p(crops.data_ptr()),
                                                     C.c_void_p(org.data_ptr()), C.c_void_p(offs.data_ptr()),
                                                     C.c_void_p(status.data_ptr()), self._stream()))
            if int(status.item()) & 1:
                raise ValueError("a box is empty after padding and clipping to the frame")
        return crops, org, offs

    def frame_status(self) -> int:
        """Status word of the device-side frame calls since the last query (bit 0: a box was empty after padding and
        clipping).  Synchronises the device and clears the word (vpb_frame_status)."""
        self._ensure()
        st = C.c_int32(0)
        _lib.check(_lib.lib().vpb_frame_status(self._handle, C.byref(st)))
        return int(st.value)

    def infer_frame(self, frame: torch.Tensor, bboxes, check: bool = False):
        """uint8 RGB frame [H,W,3] (CUDA) + boxes [n,4] -> (kpts f32 [n,K,3] (y, x, score) in FRAME pixels, idx i32 [n,K]):
        the whole per-person loop of VitInference.inference (easy_ViTPose/inference.py:258-272) as one enqueue, no host sync.
        A box that is empty after clipping only sets the engine's status word (frame_status()); `check=True` synchronises
        and raises ValueError like the reference does (pad_image / cv2.resize on an empty crop)."""
        frame, bb = self._check_frame(frame, bboxes)
        n = bb.shape[0]
        kp = torch.empty((n, self.num_keypoints, 3), dtype=torch.float32, device=frame.device)
        idx = torch.empty((n, self.num_keypoints), dtype=torch.int32, device=frame.device)
        if n:
            self._call_on_stream((frame, bb, kp, idx), lambda st: _lib.lib().vpb_infer_frame(
                self._handle, C.c_void_p(frame.data_ptr()), frame.shape[0], frame.shape[1], C.c_void_p(bb.data_ptr()), n,
                C.c_void_p(kp.data_ptr()), C.c_void_p(idx.data_ptr()), st))
            if check and self.frame_status() & 1:
                raise ValueError("a box is empty after padding and clipping to the frame")
        return kp, idx

    @staticmethod
    def _host_frame_args(frame: np.ndarray, bboxes: np.ndarray):
        if frame.dtype != np.uint8 or frame.ndim != 3 or frame.shape[2] != 3:
            raise ValueError("frame must be a uint8 RGB array [H,W,3]")
        bb = np.asarray(bboxes)
        if bb.dtype.kind == "f":
            bb = bb.round()
        return np.ascontiguousarray(frame), np.ascontiguousarray(bb.reshape(-1, 4), np.int32)

    def infer_frame_host(self, frame: np.ndarray, bboxes: np.ndarray):
        """HOST frame + boxes in, HOST keypoints out (vpb_infer_frame_host): H2D of the uint8 frame, the path, D2H, sync.
        More than max_batch boxes are processed in chunks."""
        self._ensure()
        frame, bb = self._host_frame_args(frame, bboxes)
        n = bb.shape[0]
        kp = np.empty((n, self.num_keypoints, 3), np.float32)
        idx = np.empty((n, self.num_keypoints), np.int32)
        with torch.cuda.device(self._device):
            for s in range(0, n, self.max_batch):
                m = min(self.max_batch, n - s)
                _lib.check_value(_lib.lib().vpb_infer_frame_host(
                    self._handle, frame.ctypes.data_as(C.c_void_p), frame.shape[0], frame.shape[1], bb[s:s + m].ctypes.data_as(C.c_void_p), m,
                    kp[s:s + m].ctypes.data_as(C.c_void_p), idx[s:s + m].ctypes.data_as(C.c_void_p), self._stream()))
        return kp, idx

    def submit_frame_host(self, frame: np.ndarray, bboxes: np.ndarray, kpts_out: np.ndarray, idx_out: np.ndarray, slot: int) -> None:
        """Asynchronous vpb_submit_frame_host (wait with wait_host(slot)): uint8 frame [H,W,3], int32 boxes [n,4] (already
        rounded), float32 kpts_out [n,K,3], int32 idx_out [n,K]; all C-contiguous, alive and unmodified until the wait."""
        self._ensure()
        n = bboxes.shape[0]
        if frame.dtype != np.uint8 or bboxes.dtype != np.int32 or kpts_out.dtype != np.float32 or idx_out.dtype != np.int32:
            raise TypeError("submit_frame_host takes a uint8 frame, int32 boxes / idx and float32 keypoints")
        if frame.ndim != 3 or frame.shape[2] != 3 or bboxes.shape != (n, 4) or kpts_out.shape != (n, self.num_keypoints, 3) \
                or idx_out.shape != (n, self.num_keypoints) or not all(a.flags.c_contiguous for a in (frame, bboxes, kpts_out, idx_out)):
            raise ValueError("submit_frame_host: wrong shapes or non-contiguous arrays")
        with torch.cuda.device(self._device):
            _lib.check_value(_lib.lib().vpb_submit_frame_host(
                self._handle, frame.ctypes.data_as(C.c_void_p), frame.shape[0], frame.shape[1], bboxes.ctypes.data_as(C.c_void_p), n,
                kpts_out.ctypes.data_as(C.c_void_p), idx_out.ctypes.data_as(C.c_void_p), int(slot)))

    def wait_host(self, slot: int) -> None:
        _lib.check(_lib.lib().vpb_wait_host(self._handle, int(slot)))

    def kernel_launches(self, batch: int) -> int:
        self._ensure()
        return int(_lib.lib().vpb_kernel_launches(self._handle, batch))

    def set_option(self, name: str, value: int) -> None:
        self._ensure()
        _lib.check(_lib.lib().vpb_set_option(self._handle, name.encode(), int(value)))

    def profile_collect(self) -> dict:
        """{kernel class: (total ms, launches)} since the last call; needs set_option('profile', 1)."""
        self._ensure()
        L = _lib.lib()
        n = L.vpb_profile_classes()
        ms = (C.c_float * n)()
        cnt = (C.c_int32 * n)()
        _lib.check(L.vpb_profile_collect(self._handle, ms, cnt))
        return {L.vpb_profile_class_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def read_buffer(self, name: str, shape, dtype) -> torch.Tensor:
        """Debug: synchronous copy of an internal activation buffer (see vpb_read_buffer)."""
        self._ensure()
        tdtype = torch.bfloat16 if dtype == "bf16" else torch.float32
        out = torch.empty(tuple(shape), dtype=tdtype)
        _lib.check(_lib.lib().vpb_read_buffer(self._handle, name.encode(), C.c_void_p(out.data_ptr()), out.numel() * out.element_size()))
        return out

    def __del__(self):
        try:
            if self._handle:
                _lib.lib().vpb_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass
