"""CPU oracle for the ViTPose crop hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of what the reference computes on the path
    crops [B,3,256,192] -> ViT backbone -> TopdownHeatmapSimpleHead -> heatmaps [B,K,64,48]
    -> argmax + DARK/UDP refine -> keypoints [B,K,3]
It exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg have something independent to check (or time) the CUDA
path against.  Nothing under easy_vitpose_b200/ imports it; the product path
has no CPU fallback.

Parity status: PINNED.  The reference holds no tests or golden vectors of its
own, so the oracle is pinned against outputs of the reference
itself: oracle/make_golden.py imports the reference modules from
/root/reference, runs them on seeded weights/inputs and stores the results in
tests/golden/*.npz; tests/test_oracle_golden.py checks this file against them.

Every function cites the reference lines (relative to /root/reference/) whose
arithmetic it restates.  The restatement is in GEMM / stencil form
(SURVEY.md section 9), not a transliteration of the torch modules.
"""
from __future__ import annotations

import math

import numpy as np

try:  # exact erf for GELU; scipy ships in the image
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover - scipy is present in the image
    _erf = np.vectorize(math.erf, otypes=[np.float64])

IMG_H, IMG_W = 256, 192          # easy_ViTPose/configs/ViTPose_common.py:30 (image_size=[192,256] is W,H)
PATCH = 16                       # ViTPose_common.py:71
PATCH_PAD = 2                    # backbone/vit.py:222: padding = 4 + 2*(ratio//2-1), ratio=1
GRID_H, GRID_W = 16, 12          # (256+4-16)//16+1, (192+4-16)//16+1
TOKENS = GRID_H * GRID_W         # 192
HM_H, HM_W = 64, 48              # ViTPose_common.py:31 heatmap_size=[48,64]
LN_EPS = 1e-6                    # backbone/vit.py:274
BN_EPS = 1e-5                    # torch.nn.BatchNorm2d default, head/topdown_heatmap_simple_head.py:316

# (embed_dim, depth, heads): ViTPose_common.py:72-74,105-107,138-140,171-173
MODEL_DIMS = {
    "s": (384, 12, 12),
    "b": (768, 12, 12),
    "l": (1024, 24, 16),
    "h": (1280, 32, 16),
}


# --------------------------------------------------------------------------------------
# seeded weights that every side (reference, oracle, CUDA engine) can regenerate
# --------------------------------------------------------------------------------------
def make_state_dict(embed_dim: int, depth: int, num_keypoints: int, seed: int,
                    deconv_filters: int = 256, peaky: float = 1.0, bumps: bool = False) -> dict[str, np.ndarray]:
    """Deterministic float32 weights under the reference's state_dict key names.

    Key/shape contract: SURVEY.md section 8b (probed from ViTPose(cfg).state_dict()).
    np.random.RandomState is a frozen stream, so the GPU box regenerates the same
    numbers without the reference being present.  BN running stats are randomised
    so the eval-mode fold is exercised; `peaky` scales the random part of the final 1x1 conv;
    `bumps=True` adds the signal pathway of _add_bump_pathway (one clear peak per keypoint).
    """
    rs = np.random.RandomState(seed)
    D, F = embed_dim, deconv_filters

    def nrm(*shape, std=1.0, mean=0.0):
        return (rs.standard_normal(shape) * std + mean).astype(np.float32)

    sd: dict[str, np.ndarray] = {}
    sd["backbone.pos_embed"] = nrm(1, TOKENS + 1, D, std=0.02)
    sd["backbone.patch_embed.proj.weight"] = nrm(D, 3, PATCH, PATCH, std=0.03)
    sd["backbone.patch_embed.proj.bias"] = nrm(D, std=0.02)
    for i in range(depth):
        p = f"backbone.blocks.{i}."
        sd[p + "norm1.weight"] = nrm(D, std=0.05, mean=1.0)
        sd[p + "norm1.bias"] = nrm(D, std=0.02)
        sd[p + "attn.qkv.weight"] = nrm(3 * D, D, std=0.04)
        sd[p + "attn.qkv.bias"] = nrm(3 * D, std=0.02)
        sd[p + "attn.proj.weight"] = nrm(D, D, std=0.02)
        sd[p + "attn.proj.bias"] = nrm(D, std=0.02)
        sd[p + "norm2.weight"] = nrm(D, std=0.05, mean=1.0)
        sd[p + "norm2.bias"] = nrm(D, std=0.02)
        sd[p + "mlp.fc1.weight"] = nrm(4 * D, D, std=0.03)
        sd[p + "mlp.fc1.bias"] = nrm(4 * D, std=0.02)
        sd[p + "mlp.fc2.weight"] = nrm(D, 4 * D, std=0.02)
        sd[p + "mlp.fc2.bias"] = nrm(D, std=0.02)
    sd["backbone.last_norm.weight"] = nrm(D, std=0.05, mean=1.0)
    sd["backbone.last_norm.bias"] = nrm(D, std=0.02)
    cin = D
    for li in (0, 3):
        sd[f"keypoint_head.deconv_layers.{li}.weight"] = nrm(cin, F, 4, 4, std=(0.15 if bumps else 1.0) / math.sqrt(cin))
        b = f"keypoint_head.deconv_layers.{li + 1}."
        sd[b + "weight"] = nrm(F, std=0.1, mean=1.0)
        sd[b + "bias"] = nrm(F, std=0.1)
        sd[b + "running_mean"] = nrm(F, std=0.1)
        sd[b + "running_var"] = (rs.uniform(0.5, 1.5, size=(F,))).astype(np.float32)
        sd[b + "num_batches_tracked"] = np.array(7, dtype=np.int64)
        cin = F
    sd["keypoint_head.final_layer.weight"] = nrm(num_keypoints, F, 1, 1, std=0.02 * peaky)
    sd["keypoint_head.final_layer.bias"] = nrm(num_keypoints, std=0.01)
    if bumps:
        _add_bump_pathway(sd, rs, D, F, num_keypoints)
    return sd


def _add_bump_pathway(sd: dict, rs, D: int, F: int, K: int) -> None:
    """Superimpose a 'trained-like' signal path on the random weights so heatmaps carry one clear
    peak per keypoint (random weights alone give noise fields whose argmax / Taylor step are
    ill-conditioned, which makes keypoint-pixel error meaningless):
      keypoint k owns token t_k and channel c_k = k % min(D, F).  pos_embed makes channel c_k hot at
      token t_k; both deconvs carry channel c_k -> c_k through a positive 4x4 bump kernel; the 1x1
      conv reads channel c_k into heatmap k.  Every other weight stays random, so all the arithmetic
      of the path still contributes (as noise) to the result.
    """
    kern = np.outer([1.0, 2.0, 2.0, 1.0], [1.0, 2.0, 2.0, 1.0]).astype(np.float32) / 4.0
    nch = min(D, F)
    depth = sum(1 for k in sd if k.endswith(".norm1.weight"))
    amp = 12.0 * depth / 12.0          # the residual stream's random part grows with depth; keep the bump on top
    for k in range(K):
        c = k % nch
        t = int(rs.randint(0, TOKENS))
        sd["backbone.pos_embed"][0, 1 + t, c] += np.float32(amp)
        sd["keypoint_head.deconv_layers.0.weight"][c, c] += kern * np.float32(1.5)
        sd["keypoint_head.deconv_layers.3.weight"][c, c] += kern * np.float32(1.0)
        sd["keypoint_head.final_layer.weight"][k, c, 0, 0] += np.float32(0.03)


def add_outliers(sd: dict, seed: int, stream_channels: int = 4, stream_value: float = 100.0, gelu_value: float = 13.0) -> dict:
    """Real-ViT-like outliers on top of make_state_dict (in place, returns sd):
      * `stream_channels` residual-stream channels sit at about +-`stream_value` on EVERY token (pos_embed), the "massive
        activation" channels of trained ViTs: LayerNorm statistics are then dominated by a handful of channels and the bf16
        operand of the qkv / fc1 GEMMs carries values two orders of magnitude apart;
      * eight fc1 biases per block at +-`gelu_value` and +-2*`gelu_value`: pre-GELU activations beyond the range the
        engine's tanh-form GELU was fitted on (the round-1 advisor finding: the unclamped fit flipped sign at |x| ~ 11).
    Seeded separately from the weights so that fixtures made before this existed do not change."""
    rs = np.random.RandomState(seed)
    D = sd["backbone.pos_embed"].shape[-1]
    depth = sum(1 for k in sd if k.endswith(".norm1.weight"))
    ch = rs.choice(np.arange(300, D), size=stream_channels, replace=False)      # away from the bump pathway's channels (< 256)
    # the outliers raise every row's standard deviation from ~1 to ~stream_value * sqrt(channels / D): lift the bump
    # pathway's pos_embed entries (the only ones above 5) by the same factor so the heatmaps keep one clear peak per keypoint
    pe = sd["backbone.pos_embed"]
    pe[np.abs(pe) > 5.0] *= np.float32(max(1.0, stream_value * math.sqrt(stream_channels / D)))
    signs = np.where(np.arange(stream_channels) % 2 == 0, 1.0, -1.0) * rs.uniform(0.8, 1.2, size=stream_channels)
    sd["backbone.pos_embed"][0, 1:, ch] += (signs * stream_value).astype(np.float32)[:, None]
    for i in range(depth):
        b = sd[f"backbone.blocks.{i}.mlp.fc1.bias"]
        idx = rs.choice(b.shape[0], size=8, replace=False)
        b[idx] = np.array([1, -1, 2, -2, 1, -1, 2, -2], np.float32) * np.float32(gelu_value)
    return sd


def make_crops(batch: int, seed: int) -> np.ndarray:
    """Synthetic normalised crops ~N(0,1), the distribution of (img/255-MEAN)/STD
    (easy_ViTPose/inference.py:314-318)."""
    return np.random.RandomState(seed).standard_normal((batch, 3, IMG_H, IMG_W)).astype(np.float32)


# --------------------------------------------------------------------------------------
# backbone
# --------------------------------------------------------------------------------------
def patch_rows(x: np.ndarray) -> np.ndarray:
    """im2col of Conv2d(3->D, k16, s16, p2): A[b, t, c*256+ky*16+kx] = x[b,c,16py-2+ky,16px-2+kx].

    backbone/vit.py:222 (conv geometry) and :224-228 (flatten(2).transpose).  Pixel rows
    254-255 / columns 190-191 are never read; the first 2 rows/cols of patch 0 are zero.
    """
    B = x.shape[0]
    xp = np.zeros((B, 3, IMG_H + 2 * PATCH_PAD, IMG_W + 2 * PATCH_PAD), np.float32)
    xp[:, :, PATCH_PAD:PATCH_PAD + IMG_H, PATCH_PAD:PATCH_PAD + IMG_W] = x
    xp = xp[:, :, :GRID_H * PATCH, :GRID_W * PATCH]
    a = xp.reshape(B, 3, GRID_H, PATCH, GRID_W, PATCH).transpose(0, 2, 4, 1, 3, 5)
    return np.ascontiguousarray(a).reshape(B, TOKENS, 3 * PATCH * PATCH)


def layer_norm(x: np.ndarray, w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """nn.LayerNorm(eps=1e-6), biased variance (backbone/vit.py:190,198,274,304)."""
    mu = x.mean(-1, keepdims=True, dtype=np.float32)
    xc = x - mu
    var = (xc * xc).mean(-1, keepdims=True, dtype=np.float32)
    return (xc / np.sqrt(var + np.float32(LN_EPS))) * w + b


def gelu_erf(x: np.ndarray) -> np.ndarray:
    """nn.GELU() default = exact erf form (backbone/vit.py:127,132)."""
    return (0.5 * x * (1.0 + _erf(x.astype(np.float64) / math.sqrt(2.0)))).astype(np.float32)


def attention(x: np.ndarray, sd: dict, p: str, heads: int) -> np.ndarray:
    """Attention.forward (backbone/vit.py:164-180): qkv rows are q|k|v, each head-major;
    q is scaled by hd^-0.5 BEFORE the QK^T product (:170)."""
    B, T, D = x.shape
    hd = D // heads
    qkv = x @ sd[p + "attn.qkv.weight"].T + sd[p + "attn.qkv.bias"]
    qkv = qkv.reshape(B, T, 3, heads, hd)
    q = qkv[:, :, 0].transpose(0, 2, 1, 3) * np.float32(hd ** -0.5)
    k = qkv[:, :, 1].transpose(0, 2, 1, 3)
    v = qkv[:, :, 2].transpose(0, 2, 1, 3)
    s = q @ k.transpose(0, 1, 3, 2)                       # [B,h,T,T]
    s = s - s.max(-1, keepdims=True)
    e = np.exp(s)
    pr = e / e.sum(-1, keepdims=True, dtype=np.float32)
    o = (pr @ v).transpose(0, 2, 1, 3).reshape(B, T, D)
    return o @ sd[p + "attn.proj.weight"].T + sd[p + "attn.proj.bias"]


def mlp(x: np.ndarray, sd: dict, p: str) -> np.ndarray:
    """Mlp.forward (backbone/vit.py:136-141); dropout p=0 is the identity."""
    h = gelu_erf(x @ sd[p + "mlp.fc1.weight"].T + sd[p + "mlp.fc1.bias"])
    return h @ sd[p + "mlp.fc2.weight"].T + sd[p + "mlp.fc2.bias"]


def backbone_tokens(x: np.ndarray, sd: dict, depth: int, heads: int) -> np.ndarray:
    """ViT.forward up to last_norm, token-major [B,192,D] (backbone/vit.py:375-387).
    DropPath is the identity in eval (:197, :29-30)."""
    D = sd["backbone.pos_embed"].shape[-1]
    w = sd["backbone.patch_embed.proj.weight"].reshape(D, -1)
    pos = sd["backbone.pos_embed"]
    tok = patch_rows(x) @ w.T + sd["backbone.patch_embed.proj.bias"]
    tok = tok + pos[:, 1:] + pos[:, :1]                   # vit.py:382
    for i in range(depth):
        p = f"backbone.blocks.{i}."
        tok = tok + attention(layer_norm(tok, sd[p + "norm1.weight"], sd[p + "norm1.bias"]), sd, p, heads)
        tok = tok + mlp(layer_norm(tok, sd[p + "norm2.weight"], sd[p + "norm2.bias"]), sd, p)
    return layer_norm(tok, sd["backbone.last_norm.weight"], sd["backbone.last_norm.bias"])


# --------------------------------------------------------------------------------------
# head
# --------------------------------------------------------------------------------------
# ConvTranspose2d(k4,s2,p1): output row 2m+py takes (kernel row ky, input row m+dy) for
# (ky,dy) in DECONV_TAPS[py]; same along x.  head/topdown_heatmap_simple_head.py:305-313
# with _get_deconv_cfg (topdown_heatmap_base_head.py:105-120) giving padding=1, output_padding=0.
DECONV_TAPS = {0: ((1, 0), (3, -1)), 1: ((0, 1), (2, 0))}


def fold_bn(sd: dict, bn_prefix: str) -> tuple[np.ndarray, np.ndarray]:
    """Eval-mode BatchNorm2d as per-channel scale/shift (simple_head.py:316)."""
    s = sd[bn_prefix + "weight"] / np.sqrt(sd[bn_prefix + "running_var"] + np.float32(BN_EPS))
    return s.astype(np.float32), (sd[bn_prefix + "bias"] - sd[bn_prefix + "running_mean"] * s).astype(np.float32)


def deconv_bn_relu(x: np.ndarray, w: np.ndarray, scale: np.ndarray, shift: np.ndarray) -> np.ndarray:
    """x [B,H,W,Cin] NHWC, w [Cin,Cout,4,4] -> relu(bn(deconv(x))) [B,2H,2W,Cout] via the four
    sub-pixel phase GEMMs with K = 4*Cin."""
    B, H, W, Cin = x.shape
    Cout = w.shape[1]
    xp = np.zeros((B, H + 2, W + 2, Cin), np.float32)
    xp[:, 1:-1, 1:-1] = x
    out = np.empty((B, 2 * H, 2 * W, Cout), np.float32)
    for py in (0, 1):
        for px in (0, 1):
            acc = np.zeros((B * H * W, Cout), np.float32)
            for ky, dy in DECONV_TAPS[py]:
                for kx, dx in DECONV_TAPS[px]:
                    a = xp[:, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W].reshape(B * H * W, Cin)
                    acc += a @ w[:, :, ky, kx]
            acc = acc * scale + shift
            out[:, py::2, px::2] = np.maximum(acc, 0.0).reshape(B, H, W, Cout)
    return out


def head_heatmaps(tokens: np.ndarray, sd: dict) -> np.ndarray:
    """TopdownHeatmapSimpleHead.forward (simple_head.py:188-193) on token-major features:
    2x (deconv4x4s2 -> BN -> ReLU) then Conv2d 1x1 (+bias) (:124-129).  Returns [B,K,64,48]."""
    B, T, D = tokens.shape
    x = tokens.reshape(B, GRID_H, GRID_W, D)
    for li in (0, 3):
        s, t = fold_bn(sd, f"keypoint_head.deconv_layers.{li + 1}.")
        x = deconv_bn_relu(x, sd[f"keypoint_head.deconv_layers.{li}.weight"], s, t)
    wf = sd["keypoint_head.final_layer.weight"][:, :, 0, 0]
    hm = x.reshape(-1, x.shape[-1]) @ wf.T + sd["keypoint_head.final_layer.bias"]
    return np.ascontiguousarray(hm.reshape(B, HM_H, HM_W, -1).transpose(0, 3, 1, 2)).astype(np.float32)


def forward_heatmaps(x: np.ndarray, sd: dict, depth: int, heads: int) -> np.ndarray:
    """ViTPose.forward (vit_models/model.py:23-24)."""
    return head_heatmaps(backbone_tokens(x, sd, depth, heads), sd)


# --------------------------------------------------------------------------------------
# decode: argmax + DARK/UDP Taylor refine + UDP map to crop pixels
# --------------------------------------------------------------------------------------
# cv2.getGaussianKernel(k, sigma <= 0) does not evaluate the formula for k <= 9: it returns fixed tables (OpenCV 4.13:
# small_gaussian_tab).  Pinned against cv2 itself in oracle/make_golden_modes_small.py.
_SMALL_GAUSSIAN_TAPS = {
    1: [1.0],
    3: [0.25, 0.5, 0.25],
    5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
    7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125],
    9: [4 / 256, 13 / 256, 30 / 256, 51 / 256, 60 / 256, 51 / 256, 30 / 256, 13 / 256, 4 / 256],
}


def gaussian_taps(ksize: int = 11) -> np.ndarray:
    """cv2.getGaussianKernel(ksize, sigma<=0): sigma = 0.3*((ksize-1)*0.5-1)+0.8 (=2.0 for 11),
    coefficients exp(-(i-c)^2/(2 sigma^2)) normalised to sum 1, held as float32 for a
    CV_32F image (fixed tables for ksize <= 9).  Called from vit_utils/top_down_eval.py:385 with kernel=11."""
    if ksize in _SMALL_GAUSSIAN_TAPS:
        return np.asarray(_SMALL_GAUSSIAN_TAPS[ksize], np.float32)
    sigma = 0.3 * ((ksize - 1) * 0.5 - 1.0) + 0.8
    i = np.arange(ksize, dtype=np.float64) - (ksize - 1) / 2.0
    k = np.exp(-(i * i) / (2.0 * sigma * sigma))
    return (k / k.sum()).astype(np.float32)


def _reflect101(i: np.ndarray, n: int) -> np.ndarray:
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def _fma32(a: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    """float32 fused multiply-add: the double product of two floats is exact and the double sum
    is correct to well below half a float ulp, so rounding once to float32 reproduces fmaf."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def row_pass(cols: "list[np.ndarray]", taps: np.ndarray) -> np.ndarray:
    """Row pass of cv2's separable float32 filter over the 2r+1 shifted sample arrays `cols` (cols[j] = the samples at offset
    j - r).  Kernels of 7 taps and more accumulate left to right, acc = 0; acc = fmaf(k[j], x[j], acc).  Kernels of 3 and 5
    taps take cv2's small-kernel path, which sums symmetrically and starts with the inner pair:
    acc = (x[-1] + x[+1]) * k1; acc = fmaf(k0, x[0], acc); acc = fmaf(k2, x[-2] + x[+2], acc)   (found by matching
    cv2.GaussianBlur bit for bit; oracle/make_golden_modes_small.py re-checks it on whole maps)."""
    n = len(taps)
    r = (n - 1) // 2
    if n in (3, 5):
        acc = ((cols[r - 1] + cols[r + 1]).astype(np.float32) * taps[r + 1]).astype(np.float32)
        acc = _fma32(np.broadcast_to(taps[r], acc.shape), cols[r], acc)
        if n == 5:
            acc = _fma32(np.broadcast_to(taps[r + 2], acc.shape), (cols[r - 2] + cols[r + 2]).astype(np.float32), acc)
        return acc
    acc = np.zeros(np.shape(cols[0]), np.float32)
    for j in range(n):
        acc = _fma32(np.broadcast_to(taps[j], acc.shape), cols[j], acc)
    return acc


def blur_at(h: np.ndarray, xs: np.ndarray, ys: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """Value of cv2.GaussianBlur(h, (11,11), 0) (float32, BORDER_REFLECT_101 = cv2 default) at the
    integer points (xs, ys) of map h [H,W].  Only the points the Taylor stencil consumes are
    evaluated.

    Accumulation order is the one cv2 4.13's separable float filter uses, found by matching
    cv2.GaussianBlur bit for bit on random maps (oracle/make_golden.py re-checks it):
      row pass    acc = 0; for j = 0..10 (left to right): acc = fmaf(k[j], x[c-5+j], acc)     (3 / 5 taps: see row_pass)
      column pass acc = k[5]*r[y]; for d = 1..5: acc = fmaf(k[5+d], r[y+d] + r[y-d], acc)
    """
    H, W = h.shape
    r = (len(taps) - 1) // 2
    off = np.arange(-r, r + 1)
    out = np.empty(len(xs), np.float32)
    for n, (x, y) in enumerate(zip(xs, ys)):
        rows = _reflect101(y + off, H)
        cols = _reflect101(x + off, W)
        win = h[np.ix_(rows, cols)].astype(np.float32)          # [2r+1 rows, 2r+1 cols]
        rowpass = row_pass([win[:, j] for j in range(len(off))], taps)
        acc = np.float32(taps[r] * rowpass[r])
        for d in range(1, r + 1):
            pair = np.float32(rowpass[r + d] + rowpass[r - d])
            acc = _fma32(np.asarray(taps[r + d]), np.asarray(pair), np.asarray(acc))[()]
        out[n] = acc
    return out


def argmax_first(h: np.ndarray) -> tuple[int, np.float32]:
    """np.argmax semantics: first index of the maximum (top_down_eval.py:106-107)."""
    flat = h.reshape(-1)
    idx = int(np.argmax(flat))
    return idx, flat[idx]


def decode_maps(heatmaps: np.ndarray, org_wh: np.ndarray, wrap: str = "crop",
                ksize: int = 11) -> tuple[np.ndarray, np.ndarray]:
    """heatmaps [N,K,H,W] f32, org_wh [N,2] int (crop width,height) ->
    (kpts [N,K,3] f32 rows (y, x, score) in crop pixels, idx [N,K] int32 flat argmax).

    Restates, for the branch VitInference takes (unbiased=True, use_udp=True, GaussianHeatmap):
      _get_max_preds            vit_utils/top_down_eval.py:82-114
      post_dark_udp(kernel=11)  vit_utils/top_down_eval.py:354-415
      transform_preds(use_udp)  vit_utils/post_processing/post_transforms.py:183-192
      VitInference.postprocess  easy_ViTPose/inference.py:187-205 (centre = org//2, (y,x,score) order)

    `wrap` selects what "the previous map" means for the max<=0 sentinel quirk:
    the reference's flat gather underflows into the previous map of the same call and wraps from
    the first map to the last one.  "crop": a call holds one crop (VitInference, N=1);
    "batch": one call holds all N crops (keypoints_from_heatmaps on an [N,K,H,W] array).
    """
    N, K, H, W = heatmaps.shape
    taps = gaussian_taps(ksize)
    eps = np.float64(np.finfo(np.float32).eps)
    kpts = np.empty((N, K, 3), np.float32)
    idxs = np.empty((N, K), np.int32)

    def logblur(n, k, xs, ys):
        g = blur_at(heatmaps[n, k], np.asarray(xs), np.asarray(ys), taps)
        return np.log(np.clip(g, np.float32(1e-3), np.float32(50.0))).astype(np.float32)

    for n in range(N):
        ow, oh = int(org_wh[n, 0]), int(org_wh[n, 1])
        for k in range(K):
            idx, mx = argmax_first(heatmaps[n, k])
            idxs[n, k] = idx
            if mx > 0:
                x, y = idx % W, idx // W
                cx = lambda v: min(max(v, 0), W - 1)   # np.pad(mode='edge') (:389-391)
                cy = lambda v: min(max(v, 0), H - 1)
                pts = [(cx(x), cy(y)), (cx(x + 1), cy(y)), (cx(x), cy(y + 1)), (cx(x + 1), cy(y + 1)),
                       (cx(x - 1), cy(y)), (cx(x), cy(y - 1)), (cx(x - 1), cy(y - 1))]
                l = logblur(n, k, [p[0] for p in pts], [p[1] for p in pts])
                i_, ix1, iy1, ix1y1, ix1_, iy1_, ix1_y1_ = l
            else:
                # (-1,-1) sentinel (:113): padded flat index is 0 -> four reads hit the top-left pad
                # corner of this map, three reads underflow into the previous map's bottom rows.
                x, y = -1, -1
                if wrap == "crop":
                    pn, pk = n, (k - 1) % K
                else:
                    f = (n * K + k - 1) % (N * K)
                    pn, pk = divmod(f, K)
                c = logblur(n, k, [0], [0])[0]
                lp = logblur(pn, pk, [W - 1, 0], [H - 1, H - 1])
                i_ = ix1 = iy1 = ix1y1 = c
                ix1_y1_ = lp[0]            # index-W-3 -> padded (H, W+1) of previous map
                ix1_ = lp[0]               # index-1   -> padded (H+1, W+1)
                iy1_ = lp[1]               # index-W-2 -> padded (H+1, 0)
            f32 = np.float32
            dx = f32(0.5) * (ix1 - ix1_)
            dy = f32(0.5) * (iy1 - iy1_)
            dxx = ix1 - f32(2) * i_ + ix1_
            dyy = iy1 - f32(2) * i_ + iy1_
            dxy = f32(0.5) * (ix1y1 - ix1 - iy1 + i_ + i_ - ix1_ - iy1_ + ix1_y1_)
            hes = np.array([[dxx, dxy], [dxy, dyy]], np.float64) + eps * np.eye(2)   # :413 (float64)
            hinv = np.linalg.inv(hes)
            off = hinv @ np.array([dx, dy], np.float64)
            xr = np.float32(np.float32(x) - off[0])                                    # :414 (f32 -= f64)
            yr = np.float32(np.float32(y) - off[1])
            # transform_preds(use_udp=True) with scale=(ow,oh), centre=(ow//2, oh//2); numpy promotes
            # the float32 coords with python/int64 scalars to float64, result stored into float32.
            X = np.float32(np.float64(xr) * (ow / (W - 1.0)) + (ow // 2) - ow * 0.5)
            Y = np.float32(np.float64(yr) * (oh / (H - 1.0)) + (oh // 2) - oh * 0.5)
            kpts[n, k] = (Y, X, mx)
    return kpts, idxs


def infer_crops(x: np.ndarray, org_wh: np.ndarray, sd: dict, depth: int, heads: int):
    """Whole path for a batch of normalised crops: heatmaps, keypoints (y,x,score), argmax."""
    hm = forward_heatmaps(x, sd, depth, heads)
    kpts, idx = decode_maps(hm, org_wh, wrap="crop")
    return hm, kpts, idx


# --------------------------------------------------------------------------------------
# synthetic heatmaps for decode parity (edge cases the reference's decode has to survive)
# --------------------------------------------------------------------------------------
def make_decode_maps(n: int, k: int, seed: int) -> np.ndarray:
    """[n,k,64,48] float32 maps cycling through: noisy Gaussian blobs (sigma 1..3, amplitude
    0.05..1, centres up to 2 px outside the map), all-negative maps (the (-1,-1) sentinel),
    exact two-way ties (first index must win), all-zero maps, corner peaks, pure noise and
    sub-1e-3 peaks (flattened by the clip in top_down_eval.py:386)."""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:HM_H, 0:HM_W].astype(np.float32)
    out = np.empty((n, k, HM_H, HM_W), np.float32)
    for i in range(n):
        for j in range(k):
            kind = (i * k + j) % 10
            cx, cy = rs.uniform(-2, HM_W + 1), rs.uniform(-2, HM_H + 1)
            sg, amp = rs.uniform(1.0, 3.0), rs.uniform(0.05, 1.0)
            blob = (amp * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * sg * sg))).astype(np.float32)
            noise = (rs.standard_normal((HM_H, HM_W)) * 0.01).astype(np.float32)
            if kind in (0, 1, 2, 3):
                m = blob + noise
            elif kind == 4:                                   # all negative -> sentinel
                m = -np.abs(blob + noise) - np.float32(0.05)
            elif kind == 5:                                   # exact tie, later duplicate of the max
                m = blob + noise
                a = int(np.argmax(m))
                b = (a + 1 + rs.randint(1, 2000)) % (HM_H * HM_W)
                lo, hi = min(a, b), max(a, b)
                m.reshape(-1)[lo] = m.reshape(-1)[hi] = m.max() + np.float32(0.01)
            elif kind == 6:                                   # constant zero (max == 0 -> sentinel)
                m = np.zeros((HM_H, HM_W), np.float32)
            elif kind == 7:                                   # corner / border peaks
                cxs, cys = [(0, 0), (HM_W - 1, 0), (0, HM_H - 1), (HM_W - 1, HM_H - 1)][rs.randint(4)]
                m = (amp * np.exp(-((xx - cxs) ** 2 + (yy - cys) ** 2) / (2 * sg * sg))).astype(np.float32) + noise
            elif kind == 8:                                   # pure noise (ill-conditioned Hessian)
                m = noise * np.float32(5.0)
            else:                                             # peak below the 1e-3 clip
                m = blob * np.float32(5e-4 / max(float(blob.max()), 1e-6))
            out[i, j] = m
    return out
