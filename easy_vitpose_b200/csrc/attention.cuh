// Fused multi-head attention for one ViTPose crop: T = 192 tokens, head_dim 64, everything on chip.
//
// Work item = (crop b, head h).  Per item:
//   TMA:   Q,K,V [192 x 64] bf16 boxes of the qkv activation [M, 3D]  -> 128B-swizzled smem
//   UMMA:  S = Q K^T   two M=128 tiles x N=192, fp32 in TMEM (384 columns)          (q pre-scaled)
//   SIMT:  row softmax straight out of TMEM (one thread per row, 192 threads), P -> bf16 -> smem
//          in the K-major 128B-swizzled layout the tensor core wants as an A operand
//   UMMA:  O = P V     two M=128 tiles x N=64, K=192; V is consumed as an MN-major B operand, i.e.
//          exactly the [token][dim] box TMA delivered - no transpose
//   SIMT:  O / rowsum -> bf16 -> attn_out[b*192 + t, h*64 + d]
// Rows 192..255 of the second M tile are padding: their A rows are whatever lies behind the tile in
// smem; UMMA rows are independent, so they only produce TMEM lanes nobody reads.
//
// The loads of the next item's Q/K (after S is done) and V (after O is done) are issued while the
// current item is still in softmax / epilogue, so TMA latency is off the critical path.
#pragma once
#include <cuda.h>

#include "ptx.cuh"

namespace vpb {

constexpr int ATT_T = 192;
constexpr int ATT_HD = 64;
constexpr int ATT_THREADS = 192;
constexpr int ATT_TILE_BYTES = ATT_T * ATT_HD * 2;          // 24576: one Q/K/V box
constexpr int ATT_P_ATOM = 128 * 128;                       // one 64-wide K atom of a 128-row P tile
constexpr int ATT_P_TILE = 3 * ATT_P_ATOM;                  // 49152
constexpr int ATT_SMEM = 3 * ATT_TILE_BYTES + 2 * ATT_P_TILE + 1024 + 128;
constexpr int ATT_S_COLS = 192;                             // TMEM columns per S tile
constexpr int ATT_O_COL0 = 384;                             // O tiles behind the two S tiles

struct AttnParams {
  int batch;              // crops
  int heads;
  int dim;                // D = heads * 64
  __nv_bfloat16* out;     // [batch*192, D]
  const __nv_bfloat16* qkv;  // only used by the VT_MANUAL fallback
  int v_manual;           // 0: V through TMA as MN-major operand; 1: threads build V^T (K-major) by hand
};

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_tcgen05(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT_TILE_BYTES;
  uint8_t* sP = sV + ATT_TILE_BYTES;                        // 2 tiles x 3 atoms x [128 rows x 128 B]
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(sP + 2 * ATT_P_TILE);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;
  uint64_t* bar_p = bar_qk + 3;
  uint64_t* bar_o = bar_qk + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = warp >> 2;                               // 0: rows 0..127, 1: rows 128..191
  const int quarter = warp & 3;
  const int row_in_tile = quarter * 32 + lane;
  const int t = tile * 128 + row_in_tile;                   // token (row of S / O); always < 192 here
  const int items = p.batch * p.heads;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, ATT_THREADS);
    mbar_init(bar_o, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();                                               // qkv from the previous GEMM is complete

  auto load_qk = [&](int item) {
    const int b = item / p.heads, h = item % p.heads;
    mbar_expect_tx(bar_qk, 2 * ATT_TILE_BYTES);
    tma_load_2d(sQ, &tmap_qkv, bar_qk, h * ATT_HD, b * ATT_T);
    tma_load_2d(sK, &tmap_qkv, bar_qk, p.dim + h * ATT_HD, b * ATT_T);
  };
  auto load_v = [&](int item) {
    const int b = item / p.heads, h = item % p.heads;
    mbar_expect_tx(bar_v, ATT_TILE_BYTES);
    tma_load_2d(sV, &tmap_qkv, bar_v, 2 * p.dim + h * ATT_HD, b * ATT_T);
  };

  const int first = blockIdx.x;
  if (threadIdx.x == 0 && first < items) {
    load_qk(first);
    if (!p.v_manual) load_v(first);
  }

  uint32_t phase = 0;
  for (int item = first; item < items; item += gridDim.x, phase ^= 1) {
    const int b = item / p.heads, h = item % p.heads;
    const int next = item + gridDim.x;

    if (p.v_manual) {
      // Fallback: V^T as a K-major operand: row = dim d (64 rows), K = token s; atom j = s / 64.
      // element (d, s) -> sV + j*8192 + d*128 + (((s%64)/8) ^ (d%8))*16 + (s%8)*2
      const __nv_bfloat16* vsrc = p.qkv + static_cast<size_t>(b) * ATT_T * (3 * p.dim) + 2 * p.dim + h * ATT_HD;
      for (int e = threadIdx.x; e < ATT_T * ATT_HD; e += ATT_THREADS) {
        const int s = e / ATT_HD, d = e % ATT_HD;
        const __nv_bfloat16 val = vsrc[static_cast<size_t>(s) * (3 * p.dim) + d];
        const int off = (s / 64) * 8192 + d * 128 + ((((s % 64) / 8) ^ (d % 8)) * 16) + (s % 8) * 2;
        *reinterpret_cast<__nv_bfloat16*>(sV + off) = val;
      }
      fence_proxy_async_smem();
    }

    // ---------------------------------------------------------------- S = Q K^T
    if (threadIdx.x == 0) {
      mbar_wait(bar_qk, phase);
      tc_fence_after_sync();
      constexpr uint32_t idesc = umma_idesc_bf16(128, ATT_T);
      const uint64_t kdesc = umma_desc_sw128(smem_u32(sK), 1024);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ) + mt * 128 * 128, 1024);
#pragma unroll
        for (int k = 0; k < ATT_HD / 16; ++k)
          umma_bf16(tmem_base + mt * ATT_S_COLS, qdesc + 2 * k, kdesc + 2 * k, idesc, k != 0);
      }
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, phase);
    tc_fence_after_sync();
    if (threadIdx.x == 0 && next < items) load_qk(next);    // Q/K smem is free again

    // ---------------------------------------------------------------- softmax rows -> P (bf16, smem)
    const uint32_t s_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + tile * ATT_S_COLS;
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < ATT_T; c += 32) {
      uint32_t r[32];
      tmem_ld32(s_addr + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
    }
    const float mscaled = mx * 1.4426950408889634f;
    float sum = 0.0f;
    uint8_t* prow = sP + tile * ATT_P_TILE + row_in_tile * 128;
    const int sw = row_in_tile & 7;
#pragma unroll 1
    for (int c = 0; c < ATT_T; c += 32) {
      uint32_t r[32];
      tmem_ld32(s_addr + c, r);
      tmem_ld_wait();
      float e[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        e[j] = ex2_approx(fmaf(__uint_as_float(r[j]), 1.4426950408889634f, -mscaled));
        sum += e[j];
      }
      uint8_t* patom = prow + (c >> 6) * ATT_P_ATOM;          // which 64-wide K atom
      const int chunk0 = (c & 63) >> 3;                       // first 16-byte chunk inside the atom row
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 w;
        w.x = pack_bf16(e[8 * q + 0], e[8 * q + 1]); w.y = pack_bf16(e[8 * q + 2], e[8 * q + 3]);
        w.z = pack_bf16(e[8 * q + 4], e[8 * q + 5]); w.w = pack_bf16(e[8 * q + 6], e[8 * q + 7]);
        *reinterpret_cast<uint4*>(patom + (((chunk0 + q) ^ sw) << 4)) = w;
      }
    }
    fence_proxy_async_smem();          // P was written through the generic proxy, UMMA reads via async proxy
    tc_fence_before_sync();
    mbar_arrive(bar_p);

    // ---------------------------------------------------------------- O = P V
    if (threadIdx.x == 0) {
      mbar_wait(bar_p, phase);
      tc_fence_after_sync();
      if (!p.v_manual) { mbar_wait(bar_v, phase); tc_fence_after_sync(); }
      const uint32_t idesc = umma_idesc_bf16(128, ATT_HD, /*b_mn_major=*/!p.v_manual);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int kk = 0; kk < ATT_T / 16; ++kk) {
          const uint64_t pdesc =
              umma_desc_sw128(smem_u32(sP) + mt * ATT_P_TILE + (kk >> 2) * ATT_P_ATOM + (kk & 3) * 32, 1024);
          // MN-major V: 16 tokens = two 8-row groups of 1024 B each -> +2048 B per K step.
          // K-major V^T (fallback): atom (kk/4) of [64 rows x 128 B] = 8192 B, +32 B inside the atom.
          const uint64_t vdesc = p.v_manual
                                     ? umma_desc_sw128(smem_u32(sV) + (kk >> 2) * 8192 + (kk & 3) * 32, 1024)
                                     : umma_desc_sw128(smem_u32(sV) + kk * 2048, 1024);
          umma_bf16(tmem_base + ATT_O_COL0 + mt * ATT_HD, pdesc, vdesc, idesc, kk != 0);
        }
      }
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, phase);
    tc_fence_after_sync();
    if (threadIdx.x == 0 && next < items && !p.v_manual) load_v(next);   // V smem is free again

    // ---------------------------------------------------------------- O / rowsum -> global
    const float inv = 1.0f / sum;
    const uint32_t o_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + ATT_O_COL0 + tile * ATT_HD;
    __nv_bfloat16* orow = p.out + (static_cast<size_t>(b) * ATT_T + t) * p.dim + h * ATT_HD;
#pragma unroll
    for (int c = 0; c < ATT_HD; c += 32) {
      uint32_t r[32];
      tmem_ld32(o_addr + c, r);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 w;
        w.x = pack_bf16(__uint_as_float(r[8 * q + 0]) * inv, __uint_as_float(r[8 * q + 1]) * inv);
        w.y = pack_bf16(__uint_as_float(r[8 * q + 2]) * inv, __uint_as_float(r[8 * q + 3]) * inv);
        w.z = pack_bf16(__uint_as_float(r[8 * q + 4]) * inv, __uint_as_float(r[8 * q + 5]) * inv);
        w.w = pack_bf16(__uint_as_float(r[8 * q + 6]) * inv, __uint_as_float(r[8 * q + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + c + 8 * q) = w;
      }
    }
    // S/O TMEM and (manual mode) sV are rewritten by the next item: everyone must be done reading.
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace vpb
