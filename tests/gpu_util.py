"""Helpers for the -m gpu tests: raw C-ABI calls on torch-owned device memory."""
import ctypes as C

import torch

from easy_vitpose_b200 import _lib

EPI_BF16, EPI_BF16_GELU, EPI_BF16_RELU_UP, EPI_F32_NCHW, EPI_F32_ADD, EPI_BF16_GELU_ERF = 0, 1, 2, 4, 5, 6


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(a, w, bias, out, epi, resid=None, resid_mod=0, aux=(0, 0, 0, 0)):
    m, k = a.shape
    n = w.shape[0]
    _lib.check(_lib.lib().vpb_gemm(ptr(a), ptr(w), ptr(bias), ptr(out), m, n, k, epi, ptr(resid), resid_mod,
                                   aux[0], aux[1], aux[2], aux[3], stream()))
    torch.cuda.synchronize()


def attention(qkv, batch, heads, head_dim=64):
    out = torch.empty((batch * 192, heads * head_dim), dtype=torch.bfloat16, device=qkv.device)
    _lib.check(_lib.lib().vpb_attention(ptr(qkv), batch, heads, head_dim, ptr(out), stream()))
    torch.cuda.synchronize()
    return out


def layernorm(x, g, b, eps=1e-6):
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().vpb_layernorm(ptr(x), ptr(g), ptr(b), ptr(y), x.shape[0], x.shape[1], eps, stream()))
    torch.cuda.synchronize()
    return y
