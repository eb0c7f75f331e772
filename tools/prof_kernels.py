#!/usr/bin/env python
"""Launches the hot kernels in isolation at the BASELINE configs[1] shapes (ViT-B, 64 crops -> M = 12288) so that
`ncu --set full -k regex:<name>` captures them quickly.  Order: qkv, proj, fc1, fc2 GEMMs, attention, layernorm."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gpu_util import EPI_BF16, EPI_BF16_GELU, EPI_F32_ADD, attention, gemm, layernorm

dev = torch.device("cuda", 0)
M, D = 12288, 768
torch.manual_seed(0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
xn = (torch.randn(M, D, device=dev) * 0.5).bfloat16()
hid = (torch.randn(M, 4 * D, device=dev) * 0.5).bfloat16()
wqkv = (torch.randn(3 * D, D, device=dev) * 0.03).bfloat16()
wproj = (torch.randn(D, D, device=dev) * 0.03).bfloat16()
wfc1 = (torch.randn(4 * D, D, device=dev) * 0.03).bfloat16()
wfc2 = (torch.randn(D, 4 * D, device=dev) * 0.03).bfloat16()
b3, b1, b4 = torch.randn(3 * D, device=dev), torch.randn(D, device=dev), torch.randn(4 * D, device=dev)
qkv = torch.empty(M, 3 * D, dtype=torch.bfloat16, device=dev)
h = torch.empty(M, 4 * D, dtype=torch.bfloat16, device=dev)
x = torch.randn(M, D, device=dev)
for _ in range(reps):
    gemm(xn, wqkv, b3, qkv, EPI_BF16)
    gemm(xn, wproj, b1, x, EPI_F32_ADD)
    gemm(xn, wfc1, b4, h, EPI_BF16_GELU)
    gemm(hid, wfc2, b1, x, EPI_F32_ADD)
    attention(qkv, 64, 12, 64)
    layernorm(x, b1, b1)
print("done")
