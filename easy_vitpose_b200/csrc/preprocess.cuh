// Crop pre-processing: frame uint8 RGB [H,W,3] + int boxes [n,4]  ->  normalised crops
// f32 [n,3,256,192], canvas sizes [n,2] (w,h) and frame offsets [n,2] (y,x), one launch for all people of a frame.
// HBM/L2-bound: 589 824 B written per crop; the gather side re-reads frame pixels that stay in L2.
//
// Restates, per box, what VitInference.inference does on the CPU:
//   easy_ViTPose/inference.py:259-261   box +-10 px, clipped to the frame
//   easy_ViTPose/inference.py:264-265   crop, then pad_image(crop, 3/4): zero-pad to a 3:4 canvas (vit_utils/inference.py:41-70);
//                                       the canvas is never materialised here, out-of-crop taps read 0
//   easy_ViTPose/inference.py:314-318   pre_img: cv2.resize(.., (192,256), INTER_LINEAR) on uint8, /255, (x-MEAN)/STD in
//                                       float64, HWC -> CHW, astype(float32)
//   easy_ViTPose/inference.py:270       offset (y0 - top_pad, x0 - left_pad) that maps crop keypoints back to the frame
// cv2's uint8 bilinear resize is fixed-point; the arithmetic below is its exact integer pipeline (oracle/preproc_oracle.py
// documents and pins it): int16 coefficients rint(frac * 2048), horizontal sums in int32, vertical
// (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.  The horizontal pass clamps (index, fraction) at the
// borders, the vertical pass clamps only the row indices.  Normalisation is a 3x256 table computed in float64, so the
// crops are bit-identical to the reference's.
#pragma once
#include <cuda_bf16.h>

#include <cstdint>

#include "ptx.cuh"

namespace vpb {

constexpr int PP_W = 192, PP_H = 256, PP_ROWS = 16;     // one CTA: 192 columns x 16 output rows of one crop

struct PreprocParams {
  const uint8_t* frame;       // [fh, fw, 3] RGB, row pitch `pitch` bytes
  long long pitch;
  int fh, fw;
  const int* bboxes;          // [n,4] (x0, y0, x1, y1), already rounded to int (inference.py:253)
  int n, pad;
  float* crops;               // [n,3,256,192]
  int* org_wh;                // [n,2] canvas (w, h)
  int* offs_yx;               // [n,2] (y0 - top_pad, x0 - left_pad)
  int* status;                // bit 0 set if any box is empty after clipping (may be nullptr)
};

struct PpAxis { int i0, i1, a0, a1; };

// cv2: scale = 1 / (dst / src) in double
__device__ __forceinline__ double pp_scale(int dn, int sn) {
  return __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(dn), static_cast<double>(sn)));
}
// destination index d -> the two source indices and int16 weights over a source of `sn` samples
__device__ __forceinline__ PpAxis pp_axis(int d, double scale, int sn, bool clamp_fraction) {
  const float f = static_cast<float>(__dadd_rn(__dmul_rn(static_cast<double>(d) + 0.5, scale), -0.5));
  int s = __float2int_rd(f);
  float fr = __fsub_rn(f, static_cast<float>(s));
  if (clamp_fraction) {
    if (s < 0) { s = 0; fr = 0.f; }
    if (s >= sn - 1) { s = sn - 1; fr = 0.f; }
  }
  PpAxis a;
  a.a1 = __float2int_rn(__fmul_rn(fr, 2048.f));
  a.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fr), 2048.f));
  a.i0 = min(max(s, 0), sn - 1);
  a.i1 = min(max(s + 1, 0), sn - 1);
  return a;
}

__global__ void __launch_bounds__(PP_W) crop_resize_normalise(const PreprocParams p) {
  __shared__ float s_lut[3][256];
  __shared__ PpAxis s_ay[PP_ROWS];
  const int crop = blockIdx.x, dx = threadIdx.x, dy0 = blockIdx.y * PP_ROWS;
  for (int i = dx; i < 768; i += PP_W) {                   // inference.py:32-33 MEAN / STD, float64 as in pre_img
    const int c = i >> 8, v = i & 255;
    const double mean = c == 0 ? 0.485 : (c == 1 ? 0.456 : 0.406), stdv = c == 0 ? 0.229 : (c == 1 ? 0.224 : 0.225);
    s_lut[c][v] = static_cast<float>(__ddiv_rn(__dsub_rn(__ddiv_rn(static_cast<double>(v), 255.0), mean), stdv));
  }
  const int* bb = p.bboxes + 4 * crop;
  const int x0 = min(max(bb[0] - p.pad, 0), p.fw), x1 = min(max(bb[2] + p.pad, 0), p.fw);
  const int y0 = min(max(bb[1] - p.pad, 0), p.fh), y1 = min(max(bb[3] + p.pad, 0), p.fh);
  const int w = x1 - x0, h = y1 - y0;
  __syncthreads();
  float* out = p.crops + static_cast<size_t>(crop) * 3 * PP_H * PP_W;
  if (w <= 0 || h <= 0) {                                   // the reference raises here; flag it, emit a black crop
    if (dx == 0 && blockIdx.y == 0) {
      if (p.status) atomicOr(p.status, 1);
      p.org_wh[2 * crop] = 0; p.org_wh[2 * crop + 1] = 0;
      p.offs_yx[2 * crop] = y0; p.offs_yx[2 * crop + 1] = x0;
    }
    for (int r = 0; r < PP_ROWS; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) out[(c * PP_H + dy0 + r) * PP_W + dx] = s_lut[c][0];
    return;
  }
  // pad_image: w / h < 3 / 4  <=>  4w < 3h;  int(0.75 * h) = 3h / 4,  int(w / 0.75) = 4w / 3
  int cw = w, ch = h, left = 0, top = 0;
  if (4 * w < 3 * h) { cw = (3 * h) / 4; left = (cw - w) / 2; }
  else { ch = (4 * w) / 3; top = (ch - h) / 2; }
  if (dx == 0 && blockIdx.y == 0) {
    p.org_wh[2 * crop] = cw; p.org_wh[2 * crop + 1] = ch;
    p.offs_yx[2 * crop] = y0 - top; p.offs_yx[2 * crop + 1] = x0 - left;
  }
  if (dx < PP_ROWS) s_ay[dx] = pp_axis(dy0 + dx, pp_scale(PP_H, ch), ch, false);   // the row weights are shared by the CTA
  const PpAxis ax = pp_axis(dx, pp_scale(PP_W, cw), cw, true);
  __syncthreads();
  const int cx0 = ax.i0 - left, cx1 = ax.i1 - left;          // canvas column -> crop column
  const bool vx0 = cx0 >= 0 && cx0 < w, vx1 = cx1 >= 0 && cx1 < w;
  const uint8_t* col0 = p.frame + static_cast<size_t>(vx0 ? x0 + cx0 : 0) * 3;   // only dereferenced when valid
  const uint8_t* col1 = p.frame + static_cast<size_t>(vx1 ? x0 + cx1 : 0) * 3;
#pragma unroll 4
  for (int r = 0; r < PP_ROWS; ++r) {
    const int dy = dy0 + r;
    const PpAxis ay = s_ay[r];
    const int cy0 = ay.i0 - top, cy1 = ay.i1 - top;
    const bool vy0 = cy0 >= 0 && cy0 < h, vy1 = cy1 >= 0 && cy1 < h;
    const size_t r0 = static_cast<size_t>(vy0 ? y0 + cy0 : 0) * p.pitch, r1 = static_cast<size_t>(vy1 ? y0 + cy1 : 0) * p.pitch;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int p00 = (vy0 && vx0) ? col0[r0 + c] : 0, p01 = (vy0 && vx1) ? col1[r0 + c] : 0;
      const int p10 = (vy1 && vx0) ? col0[r1 + c] : 0, p11 = (vy1 && vx1) ? col1[r1 + c] : 0;
      const int s0 = p00 * ax.a0 + p01 * ax.a1, s1 = p10 * ax.a0 + p11 * ax.a1;
      int v = (((ay.a0 * (s0 >> 4)) >> 16) + ((ay.a1 * (s1 >> 4)) >> 16) + 2) >> 2;
      v = min(max(v, 0), 255);
      out[(c * PP_H + dy) * PP_W + dx] = s_lut[c][v];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The same pre-processing fused with the patch-embedding im2col (pointwise.cuh: patch_im2col): frame + boxes -> bf16 patch
// rows [n*192, 768] directly, so the f32 crops (589 824 B each) are never written or re-read.  Values are bf16(table[v]),
// i.e. exactly what patch_im2col produces from crop_resize_normalise's output.  One CTA = one crop x one patch row (16 image
// rows incl. the conv's 2-pixel zero border): the 16 x 192 x 3 pixels are gathered with lanes along the image row into a
// shared bf16 tile, which is then written out as 16-byte chunks of the im2col rows.
// Like patch_im2col, the launch also seeds the fp32 token stream with pos_embed + conv bias (vit.py:382).
struct FramePatchParams {
  PreprocParams pp;             // crops / status unused
  __nv_bfloat16* rows;          // [n*192, 768]
  const float4* pos_bias;       // [192*D/4]
  float4* stream;               // [n*192*D/4]
  int D;
};

__global__ void __launch_bounds__(384) frame_to_patch_rows(const FramePatchParams q) {
  constexpr int FP_PITCH = 208;                               // 2 + 192 + 14 bf16 per tile row: 16-byte aligned rows
  __shared__ uint16_t s_lut[3][256];
  __shared__ PpAxis s_ay[16];
  __shared__ PpAxis s_ax[PP_W];
  __shared__ __align__(16) uint16_t s_tile[3 * 16 * FP_PITCH];
  const PreprocParams& p = q.pp;
  const int crop = blockIdx.x, py = blockIdx.y, tid = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();                                               // the previous step may still be reading patch rows / the stream
  {
    const int per_crop4 = 192 * q.D / 4, per_cta4 = per_crop4 / 16;      // this CTA seeds 1/16 of its crop's tokens
    const float4* src = q.pos_bias + py * per_cta4;
    float4* dst = q.stream + static_cast<size_t>(crop) * per_crop4 + py * per_cta4;
    for (int j = tid; j < per_cta4; j += 384) dst[j] = __ldg(src + j);
  }
  for (int i = tid; i < 768; i += 384) {
    const int c = i >> 8, v = i & 255;
    const double mean = c == 0 ? 0.485 : (c == 1 ? 0.456 : 0.406), stdv = c == 0 ? 0.229 : (c == 1 ? 0.224 : 0.225);
    const float f = static_cast<float>(__ddiv_rn(__dsub_rn(__ddiv_rn(static_cast<double>(v), 255.0), mean), stdv));
    s_lut[c][v] = __bfloat16_as_ushort(__float2bfloat16_rn(f));
  }
  const int* bb = p.bboxes + 4 * crop;
  const int x0 = min(max(bb[0] - p.pad, 0), p.fw), x1 = min(max(bb[2] + p.pad, 0), p.fw);
  const int y0 = min(max(bb[1] - p.pad, 0), p.fh), y1 = min(max(bb[3] + p.pad, 0), p.fh);
  int w = x1 - x0, h = y1 - y0;
  const bool empty = w <= 0 || h <= 0;
  if (empty) { w = 0; h = 0; }
  int cw = max(w, 1), ch = max(h, 1), left = 0, top = 0;    // an empty box reads nothing: a black crop, as in crop_resize_normalise
  if (!empty) {
    if (4 * w < 3 * h) { cw = (3 * h) / 4; left = (cw - w) / 2; }
    else { ch = (4 * w) / 3; top = (ch - h) / 2; }
  }
  if (tid == 0 && py == 0) {
    if (empty && p.status) atomicOr(p.status, 1);
    p.org_wh[2 * crop] = empty ? 0 : cw; p.org_wh[2 * crop + 1] = empty ? 0 : ch;
    p.offs_yx[2 * crop] = y0 - top; p.offs_yx[2 * crop + 1] = x0 - left;
  }
  if (tid < PP_W) s_ax[tid] = pp_axis(tid, pp_scale(PP_W, cw), cw, true);
  else if (tid < PP_W + 16) {
    const int dy = 16 * py - 2 + (tid - PP_W);
    if (dy >= 0) s_ay[tid - PP_W] = pp_axis(dy, pp_scale(PP_H, ch), ch, false);
  }
  __syncthreads();
  // Phase 1: one (row, column) pixel per thread-iteration with consecutive lanes on consecutive output columns, so that a
  // warp's byte gathers fall into neighbouring sectors; bf16 values go to a shared tile laid out [channel][ky][2 + dx]
  // (column 0,1 and 194..207 = the conv's zero padding / alignment).
  for (int i = tid; i < 3 * 16 * 8; i += 384) {              // zero the padding columns once: 16 bf16 per (c, ky) row
    const int row = i >> 3, e = i & 7;
    s_tile[row * FP_PITCH + (e < 2 ? e : 192 + e)] = 0;
  }
  for (int i = tid; i < 16 * PP_W; i += 384) {
    const int ky = i / PP_W, dx = i % PP_W;
    const int dy = 16 * py - 2 + ky;
    uint16_t v3[3] = {0, 0, 0};
    if (dy >= 0) {
      const PpAxis ay = s_ay[ky];
      const PpAxis ax = s_ax[dx];
      const int cy0 = ay.i0 - top, cy1 = ay.i1 - top, cx0 = ax.i0 - left, cx1 = ax.i1 - left;
      const bool vy0 = cy0 >= 0 && cy0 < h, vy1 = cy1 >= 0 && cy1 < h, vx0 = cx0 >= 0 && cx0 < w, vx1 = cx1 >= 0 && cx1 < w;
      const uint8_t* r0 = p.frame + static_cast<size_t>(vy0 ? y0 + cy0 : 0) * p.pitch;
      const uint8_t* r1 = p.frame + static_cast<size_t>(vy1 ? y0 + cy1 : 0) * p.pitch;
      const int f0 = (vx0 ? x0 + cx0 : 0) * 3, f1 = (vx1 ? x0 + cx1 : 0) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int p00 = (vy0 && vx0) ? r0[f0 + c] : 0, p01 = (vy0 && vx1) ? r0[f1 + c] : 0;
        const int p10 = (vy1 && vx0) ? r1[f0 + c] : 0, p11 = (vy1 && vx1) ? r1[f1 + c] : 0;
        const int s0 = p00 * ax.a0 + p01 * ax.a1, s1 = p10 * ax.a0 + p11 * ax.a1;
        int v = (((ay.a0 * (s0 >> 4)) >> 16) + ((ay.a1 * (s1 >> 4)) >> 16) + 2) >> 2;
        v3[c] = s_lut[c][min(max(v, 0), 255)];
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) s_tile[(c * 16 + ky) * FP_PITCH + 2 + dx] = v3[c];
  }
  __syncthreads();
  // Phase 2: the im2col rows of this patch row, 16 bytes (8 kx) per store: element (px, c, ky, kx) = tile[c][ky][16 px + kx]
  for (int i = tid; i < 12 * 3 * 16 * 2; i += 384) {
    const int hf = i & 1, ky = (i >> 1) & 15, c = (i >> 5) % 3, px = i / 96;
    const uint4 v = *reinterpret_cast<const uint4*>(&s_tile[(c * 16 + ky) * FP_PITCH + 16 * px + 8 * hf]);
    *reinterpret_cast<uint4*>(q.rows + ((static_cast<size_t>(crop) * 16 + py) * 12 + px) * 768 + c * 256 + ky * 16 + 8 * hf) = v;
  }
}

}  // namespace vpb
