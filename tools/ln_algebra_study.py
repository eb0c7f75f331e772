#!/usr/bin/env python
"""CPU study for the next round: can LayerNorm be folded into the consumer GEMM?

  LN(x) W^T + b  =  rstd * (x (g*W)^T)  -  rstd * mean * colsum(g*W)  +  (beta W^T + b)

Emulates the engine's numerics in torch on the CPU (bf16 operands rounded where the engine rounds them, fp32 accumulation,
fp32 residual stream) in two variants and compares the heatmaps of both with the fp32 forward:
  A  today's path:   A operand = bf16(LN(x)),  W = bf16(W)
  B  folded path:    A operand = bf16(x),      W = bf16(g*W), row statistics applied to the fp32 accumulator
Run: python tools/ln_algebra_study.py [size=s] [crops=2]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F

from easy_vitpose_b200.synthetic import random_crops, random_state_dict
from oracle import torch_ref as T          # the fp32 forward it is compared with

torch.set_grad_enabled(False)
bf = lambda t: t.to(torch.bfloat16).to(torch.float32)


def engine_like(x, sd, depth, heads, folded):
    B = x.shape[0]
    D = sd["backbone.pos_embed"].shape[2]
    hd = D // heads
    t = F.conv2d(bf(x), bf(sd["backbone.patch_embed.proj.weight"]), None, stride=16, padding=2).flatten(2).transpose(1, 2)
    pos = sd["backbone.pos_embed"]
    t = t + (pos[:, 1:] + pos[:, :1] + sd["backbone.patch_embed.proj.bias"])

    def ln_linear(t, g, beta, W, b):
        if not folded:
            return F.linear(bf(F.layer_norm(t, (D,), g, beta, 1e-6)), bf(W), b)
        mean = t.mean(-1, keepdim=True)
        rstd = torch.rsqrt(t.var(-1, unbiased=False, keepdim=True) + 1e-6)
        Wg = bf(W * g[None, :])
        acc = F.linear(bf(t), Wg)                                   # fp32 accumulate of bf16 x bf16
        return rstd * acc - (rstd * mean) * Wg.sum(1)[None, None, :] + (F.linear(beta[None, :], W)[0] + b)

    for i in range(depth):
        p = f"backbone.blocks.{i}."
        Wq = sd[p + "attn.qkv.weight"].clone(); bq = sd[p + "attn.qkv.bias"].clone()
        Wq[:D] *= hd ** -0.5; bq[:D] *= hd ** -0.5                   # q pre-scaled in the packed weights
        qkv = bf(ln_linear(t, sd[p + "norm1.weight"], sd[p + "norm1.bias"], Wq, bq)).reshape(B, -1, 3, heads, hd).permute(2, 0, 3, 1, 4)
        s = qkv[0] @ qkv[1].transpose(-2, -1)
        pm = torch.exp(s - s.amax(-1, keepdim=True))
        o = bf((bf(pm) @ qkv[2]) / pm.sum(-1, keepdim=True)).transpose(1, 2).reshape(B, -1, D)
        t = t + F.linear(o, bf(sd[p + "attn.proj.weight"]), sd[p + "attn.proj.bias"])
        h = bf(F.gelu(ln_linear(t, sd[p + "norm2.weight"], sd[p + "norm2.bias"], sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])))
        t = t + F.linear(h, bf(sd[p + "mlp.fc2.weight"]), sd[p + "mlp.fc2.bias"])
    f = bf(F.layer_norm(t, (D,), sd["backbone.last_norm.weight"], sd["backbone.last_norm.bias"], 1e-6)).permute(0, 2, 1).reshape(B, D, 16, 12)
    for li in (0, 3):
        b = f"keypoint_head.deconv_layers.{li + 1}."
        sc = sd[b + "weight"] / torch.sqrt(sd[b + "running_var"] + 1e-5)
        w = bf(sd[f"keypoint_head.deconv_layers.{li}.weight"] * sc[None, :, None, None])
        f = bf(F.relu(F.conv_transpose2d(f, w, stride=2, padding=1) + (sd[b + "bias"] - sd[b + "running_mean"] * sc)[None, :, None, None]))
    return F.conv2d(f, bf(sd["keypoint_head.final_layer.weight"]), sd["keypoint_head.final_layer.bias"])


def main():
    size = sys.argv[1] if len(sys.argv) > 1 else "s"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dims = {"s": (384, 12, 12), "b": (768, 12, 12)}[size]
    sd = {k: torch.from_numpy(v) for k, v in random_state_dict(size, 17, seed=1).items()}
    x = torch.from_numpy(random_crops(n, 2))
    ref = T.forward(x, sd, dims[1], dims[2])
    rng = float(ref.max() - ref.min())
    for folded in (False, True):
        hm = engine_like(x, sd, dims[1], dims[2], folded)
        err = (hm - ref).abs()
        same = (hm.flatten(2).argmax(-1) == ref.flatten(2).argmax(-1)).float().mean()
        print(f"ViT-{size.upper()} {'B folded LN (bf16(x) operand)' if folded else 'A today (bf16(LN(x)) operand) '}: heatmap Linf {float(err.max()):.5f} "
              f"= {float(err.max()) / rng:.3%} of range, mean |err| {float(err.mean()):.6f}, argmax equal {float(same):.3f}")
    # how large is the row mean relative to the row std in the stream (decides how much precision bf16(x) loses vs bf16(x - mean))
    print("note: the synthetic stream is not a trained one; |mean|/std per token decides the loss and has to be measured on real checkpoints")


if __name__ == "__main__":
    main()
