"""Plain PyTorch restatement of the ViTPose forward (library kernels: cuBLAS / cuDNN / ATen)  --  TEST / BASELINE
INFRASTRUCTURE ONLY, like the rest of oracle/.

The reference's own modules cannot travel to the GPU box (Python package under /root/reference), so this file restates
ViTPose.forward with the same torch ops the reference modules call, on the same state_dict:
  PatchEmbed      F.conv2d(k16,s16,p2) + flatten/transpose             backbone/vit.py:222-228
  pos embed       x + pos[:,1:] + pos[:,:1]                            backbone/vit.py:382
  Block           LN -> qkv Linear -> q*scale @ k^T -> softmax -> @v -> proj ; LN -> fc1 -> GELU -> fc2   vit.py:136-141,164-180,202-205
  head            2 x (F.conv_transpose2d(k4,s2,p1) -> F.batch_norm(eval) -> relu) -> F.conv2d 1x1        simple_head.py:188-193,291-321
It is used (a) by tests to cross-check the numpy oracle and (b) by bench.py as the "reference torch-CUDA eager" baseline
the north star compares against (fp32 as shipped, and .to(bfloat16)).  Never imported by the product.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def forward(x: torch.Tensor, sd: dict, depth: int, heads: int) -> torch.Tensor:
    """x [B,3,256,192], sd: reference state_dict (tensors on x.device, x.dtype) -> heatmaps [B,K,64,48]."""
    B = x.shape[0]
    t = F.conv2d(x, sd["backbone.patch_embed.proj.weight"], sd["backbone.patch_embed.proj.bias"], stride=16, padding=2)
    D = t.shape[1]
    t = t.flatten(2).transpose(1, 2)
    pos = sd["backbone.pos_embed"]
    t = t + pos[:, 1:] + pos[:, :1]
    hd = D // heads
    scale = hd ** -0.5
    for i in range(depth):
        p = f"backbone.blocks.{i}."
        y = F.layer_norm(t, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
        qkv = F.linear(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, -1, 3, heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = ((q * scale) @ k.transpose(-2, -1)).softmax(dim=-1)
        y = (attn @ v).transpose(1, 2).reshape(B, -1, D)
        t = t + F.linear(y, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        y = F.layer_norm(t, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
        y = F.gelu(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
        t = t + F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    t = F.layer_norm(t, (D,), sd["backbone.last_norm.weight"], sd["backbone.last_norm.bias"], 1e-6)
    f = t.permute(0, 2, 1).reshape(B, D, 16, 12).contiguous()
    for li in (0, 3):
        f = F.conv_transpose2d(f, sd[f"keypoint_head.deconv_layers.{li}.weight"], stride=2, padding=1)
        b = f"keypoint_head.deconv_layers.{li + 1}."
        f = F.relu(F.batch_norm(f, sd[b + "running_mean"], sd[b + "running_var"], sd[b + "weight"], sd[b + "bias"], False, 0.0, 1e-5))
    return F.conv2d(f, sd["keypoint_head.final_layer.weight"], sd["keypoint_head.final_layer.bias"])


def to_device(sd_np: dict, device, dtype) -> dict:
    import numpy as np
    return {k: torch.from_numpy(np.asarray(v)).to(device=device, dtype=dtype)
            for k, v in sd_np.items() if not k.endswith("num_batches_tracked")}
