// Fused attention, head_dim 32 / 64, with the two 64-row half tiles of a pair of heads PACKED into one 128-lane pass.
//
// attention.cuh runs every (crop, head) item as two 128-row steps; the second only has 64 live rows (192 tokens), so a
// quarter of all softmax lanes idle and an item costs two full passes of the latency chain that bounds the kernel
// (S -> max -> exp -> P -> PV).  Here items are taken in PAIRS (heads 2j and 2j+1 of one crop: heads is even for every ViT)
// and a pair costs THREE steps:
//     kind 0   rows 0..127 of item A       S = Q K^T as one M=128 UMMA group          (as in attention.cuh)
//     kind 1   rows 0..127 of item B
//     kind 2   rows 128..191 of A and of B  two M=64 UMMA groups into ONE S buffer: A at TMEM lane offset 0 of every 32-lane
//              sub-partition, B at lane offset 16.
// A cta_group::1 UMMA with M = 64 keeps row m in TMEM lane 32*(m/16) + m%16 (16 lanes per sub-partition); both the
// accumulator and the A operand of a TS-MMA may sit at lane offset 16 instead of 0 as long as they use the SAME offset
// (measured on B200 by tools/experiments/umma_m64_probe.cu: accumulator and TMEM operand at offset 16 work, mixed offsets
// fault with "misaligned address").  The softmax, one thread per TMEM lane, does not care which matrix row its lane holds:
// in a packed step thread (quarter q, lane L) owns token 128 + 16q + (L & 15) of item A (L < 16) or B (L >= 16), and
// O = P V runs as two M=64 TS-MMA groups (P_A at offset 0 with V_A, P_B at offset 16 with V_B).
// 768 items: 1152 steps instead of 1536, 7.8 -> 8 per CTA instead of 10.4 -> 11.
//
// Both items of a pair stay in shared memory until their packed step, so Q, K, V live in THREE item stages (3 x 72 KB at
// head_dim 64) and the output goes from registers to global memory directly (no staging tile: 16-byte stores, every thread
// a contiguous 2*hd-byte row segment).  head_dim 80 (three stages do not fit) stays with attention.cuh.
// Warp roles, barriers and the in-place P are those of attention.cuh; see there for the pipeline description.
#pragma once
#include "attention.cuh"

namespace vpb {

template <int HD>
struct AttPackCfg {
  static_assert(HD == 32 || HD == 64, "head_dim of the packed kernel");
  static constexpr int ROW = HD * 2;                          // bytes per operand row = swizzle span (64 or 128)
  static constexpr int OPER_BYTES = ATT_T * ROW;              // one of Q / K / V: 12288 / 24576
  static constexpr int STAGE_BYTES = 3 * OPER_BYTES;
  static constexpr int STAGES = 3;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 2048 /*row sums: 4 slots x 128 rows*/ + 1024 /*align*/ + 256 /*barriers*/;
};

template <int HD, int NPOLY = 0>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_pack_tcgen05(const __grid_constant__ CUtensorMap tmap_main, const AttnParams p) {
  using Cfg = AttPackCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_sum = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);   // [4 slots = step & 3][128 lanes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_sum + 512);
  uint64_t* qk_full = bars;          // [3] Q,K of an item stage landed           (TMA -> QK issuer)
  uint64_t* v_full = bars + 3;       // [3] V landed                              (TMA -> PV issuer)
  uint64_t* s_full = bars + 6;       // [2] S complete                            (MMA commit -> softmax group step&1)
  uint64_t* p_ready = bars + 8;      // [2] row sums published                    (128 softmax threads -> epilogue)
  uint64_t* o_full = bars + 10;      // [2] O complete                            (MMA commit -> epilogue, issuers)
  uint64_t* s_free = bars + 12;      // [2] O drained                             (128 epilogue threads -> issuers)
  uint64_t* p_chunk = bars + 14;     // [2][3] P keys 64c..64c+63 published       (128 softmax threads -> PV issuer)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  // this CTA's contiguous range of steps [u0, u1): step u = kind (u % 3) of pair (u / 3); pair j = items 2j, 2j+1
  const long long steps_all = 3LL * (static_cast<long long>(p.batch) * p.heads / 2);
  const int u0 = static_cast<int>(steps_all * blockIdx.x / gridDim.x);
  const int u1 = static_cast<int>(steps_all * (blockIdx.x + 1) / gridDim.x);
  const int T = u1 - u0;
  // first item this CTA needs: item A of the first pair, unless its only step is that pair's kind 1 (every load issued must be
  // consumed: a CTA must not exit with a TMA load in flight); liA = -1 is then never used
  const int item0 = 2 * (u0 / 3) + ((u0 % 3 == 1 && T == 1) ? 1 : 0);
  const int last_u = u1 - 1;
  const int n_items = T > 0 ? 2 * (last_u / 3) + (last_u % 3 == 0 ? 0 : 1) - item0 + 1 : 0;
  const long long t_cta0 = p.dbg ? clock64() : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_main);
    for (int i = 0; i < 3; ++i) { mbar_init(&qk_full[i], 1); mbar_init(&v_full[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_free[i], 128);
      for (int c = 0; c < 3; ++c) mbar_init(&p_chunk[i * 3 + c], 128);
    }
    fence_mbar_init();
  }
  if (warp == 12) tmem_alloc(tmem_slot, ATT_TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);
  pdl_launch_dependents();
  pdl_wait();                                               // qkv from the previous GEMM is complete

  auto stage_ptr = [&](int li, int oper) { return smem + (li % 3) * Cfg::STAGE_BYTES + oper * Cfg::OPER_BYTES; };
  // local step t -> global step u = u0 + t: pair u / 3, kind u % 3; local items liA = 2 * pair - item0, liB = liA + 1 (item li
  // lives in stage li % 3, its (li / 3)-th use: barrier phase (li / 3) & 1); TMEM buffer t & 1, used n = t >> 1 times before
  constexpr uint32_t kLane16 = 16u << 16;                   // TMEM address of lane offset 16

  if (warp == 12) {
    // ------------------------------------------------------------------ issue warp 1: TMA of Q,K and S = Q K^T
    auto load_qk = [&](int li) {
      const int item = item0 + li, b = item / p.heads, h = item % p.heads;
      if (elect_one()) {
        mbar_expect_tx(&qk_full[li % 3], 2 * Cfg::OPER_BYTES);
        tma_load_2d(stage_ptr(li, 0), &tmap_main, &qk_full[li % 3], h * HD, b * ATT_T);
        tma_load_2d(stage_ptr(li, 1), &tmap_main, &qk_full[li % 3], p.dim + h * HD, b * ATT_T);
      }
      __syncwarp();
    };
    constexpr uint32_t idesc_s128 = umma_idesc_bf16(128, ATT_T);
    constexpr uint32_t idesc_s64 = umma_idesc_bf16(64, ATT_T);
    for (int li = 0; li < 3 && li < n_items; ++li) load_qk(li);
    for (int t = 0; t < T; ++t) {
      const int u = u0 + t, kind = u % 3, liA = 2 * (u / 3) - item0, liB = liA + 1, n = t >> 1;
      if (t >= 2) mbar_wait(&o_full[t & 1], (n - 1) & 1);     // PV(t-2) has consumed P: the S buffer is free (O lives elsewhere)
      if (kind != 1) mbar_wait(&qk_full[liA % 3], (liA / 3) & 1);
      if (kind != 0) mbar_wait(&qk_full[liB % 3], (liB / 3) & 1);
      tc_fence_after_sync();
      const uint32_t d = tmem_base + (t & 1) * ATT_BUF_COLS;
      if (kind < 2) {
        const int li = kind == 0 ? liA : liB;
        const uint64_t qd = umma_desc_rows<Cfg::ROW>(smem_u32(stage_ptr(li, 0)));
        const uint64_t kd = umma_desc_rows<Cfg::ROW>(smem_u32(stage_ptr(li, 1)));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_bf16(d, qd + 2 * k, kd + 2 * k, idesc_s128, k != 0);
          umma_commit(&s_full[t & 1]);
        }
      } else {
        const uint64_t qa = umma_desc_rows<Cfg::ROW>(smem_u32(stage_ptr(liA, 0)) + 128 * Cfg::ROW);
        const uint64_t ka = umma_desc_rows<Cfg::ROW>(smem_u32(stage_ptr(liA, 1)));
        const uint64_t qb = umma_desc_rows<Cfg::ROW>(smem_u32(stage_ptr(liB, 0)) + 128 * Cfg::ROW);
        const uint64_t kb = umma_desc_rows<Cfg::ROW>(smem_u32(stage_ptr(liB, 1)));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_bf16(d, qa + 2 * k, ka + 2 * k, idesc_s64, k != 0);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_bf16(d + kLane16, qb + 2 * k, kb + 2 * k, idesc_s64, k != 0);
          umma_commit(&s_full[t & 1]);
        }
      }
      __syncwarp();
      if (t >= 1 && (u - 1) % 3 == 2) {
        // step t-1 was a packed step: once its Q K^T has retired, the Q,K halves of its two item stages are free; they take
        // the next pair's item B and the item A after that (this pair's partner, item A of the next pair, is resident)
        const int pA = 2 * ((u - 1) / 3) - item0;
        mbar_wait(&s_full[(t - 1) & 1], ((t - 1) >> 1) & 1);
        if (pA + 3 < n_items) load_qk(pA + 3);
        if (pA + 4 < n_items) load_qk(pA + 4);
      }
    }
  } else if (warp == 13) {
    // ------------------------------------------------------------------ issue warp 2: TMA of V and O = P V
    auto load_v = [&](int li) {
      const int item = item0 + li, b = item / p.heads, h = item % p.heads;
      if (elect_one()) {
        mbar_expect_tx(&v_full[li % 3], Cfg::OPER_BYTES);
        tma_load_2d(stage_ptr(li, 2), &tmap_main, &v_full[li % 3], 2 * p.dim + h * HD, b * ATT_T);
      }
      __syncwarp();
    };
    constexpr uint32_t idesc_o128 = umma_idesc_bf16(128, HD, /*b_mn_major=*/true);
    constexpr uint32_t idesc_o64 = umma_idesc_bf16(64, HD, /*b_mn_major=*/true);
    for (int li = 0; li < 3 && li < n_items; ++li) load_v(li);
    for (int t = 0; t < T; ++t) {
      const int u = u0 + t, kind = u % 3, liA = 2 * (u / 3) - item0, liB = liA + 1, n = t >> 1, bf = t & 1;
      if (kind != 1) mbar_wait(&v_full[liA % 3], (liA / 3) & 1);
      if (kind != 0) mbar_wait(&v_full[liB % 3], (liB / 3) & 1);
      if (t >= 2) mbar_wait(&s_free[bf], (n - 1) & 1);        // O(t-2) has left this step parity's O columns
      const uint32_t buf = tmem_base + bf * ATT_BUF_COLS;
      const uint32_t od = tmem_base + ATT_O_SEP_COL + bf * 64;
      const uint32_t sVa = smem_u32(stage_ptr(kind == 1 ? liB : liA, 2)), sVb = smem_u32(stage_ptr(liB, 2));
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        mbar_wait(&p_chunk[bf * 3 + c], n & 1);               // keys 64c..64c+63 of P(t) are in TMEM
        tc_fence_after_sync();
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // 16 keys = 8 packed TMEM columns of P; V (MN-major): 16 tokens = two 8-row groups of the box
            const int kk = 4 * c + j;
            if (kind < 2) {
              umma_bf16_ts(od, buf + kk * 8, umma_desc_rows<Cfg::ROW>(sVa + kk * 16 * Cfg::ROW), idesc_o128, kk != 0);
            } else {
              umma_bf16_ts(od, buf + kk * 8, umma_desc_rows<Cfg::ROW>(sVa + kk * 16 * Cfg::ROW), idesc_o64, kk != 0);
              umma_bf16_ts(od + kLane16, buf + kLane16 + kk * 8, umma_desc_rows<Cfg::ROW>(sVb + kk * 16 * Cfg::ROW), idesc_o64, kk != 0);
            }
          }
          if (c == 2) umma_commit(&o_full[bf]);
        }
        __syncwarp();
      }
      if (kind == 2 && liA + 3 < n_items) {
        mbar_wait(&o_full[bf], n & 1);                       // the pair's last P V has retired: both V stages are free
        load_v(liA + 3);
        if (liA + 4 < n_items) load_v(liA + 4);
      }
    }
  } else if (warp < 8) {
    // -------------------------------------------------------------------- softmax (group A: warps 0..3, B: warps 4..7)
    // identical for the three kinds of step: a thread turns the 192 logits of ITS TMEM lane into P, whatever row that is
    const int quarter = warp & 3;
    const int grp = warp >> 2;
    const int tl = quarter * 32 + lane;
    const uint32_t buf = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + grp * ATT_BUF_COLS;
    constexpr float kLog2e = 1.4426950408889634f;
    long long w_wait = 0, w_busy = 0, c0 = 0;
    for (int t = grp; t < T; t += 2) {
      const int n = t >> 1;
      if (p.dbg) c0 = clock64();
      mbar_wait(&s_full[grp], n & 1);
      if (p.dbg) { w_wait += clock64() - c0; c0 = clock64(); }
      tc_fence_after_sync();
      uint32_t ra[32], rb[32];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 6; c += 2) {
        tmem_ld32(buf + 32 * c, ra);
        tmem_ld32(buf + 32 * c + 32, rb);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2)
          mx = fmaxf(fmaxf(mx, fmaxf(__uint_as_float(ra[j]), __uint_as_float(ra[j + 1]))), fmaxf(__uint_as_float(rb[j]), __uint_as_float(rb[j + 1])));
      }
      const float mscaled = mx * kLog2e;
      float sum = 0.0f;
      auto exp_chunk = [&](const uint32_t (&r)[32], int c) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float a0 = fmaf(__uint_as_float(r[j]), kLog2e, -mscaled), a1 = fmaf(__uint_as_float(r[j + 1]), kLog2e, -mscaled);
          const float e0 = ex2_approx(a0);
          const float e1 = (NPOLY > 0 && (((j >> 1) * NPOLY) % 16 < NPOLY)) ? ex2_poly(a1) : ex2_approx(a1);   // NPOLY of every 32
          sum += e0 + e1;
          pk[j >> 1] = pack_bf16(e0, e1);
        }
        tmem_st16(buf + 16 * c, pk);
        if (c & 1) {
          tmem_st_wait();
          tc_fence_before_sync();
          mbar_arrive(&p_chunk[grp * 3 + (c >> 1)]);
        }
      };
      tmem_ld32(buf, ra);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 6; c += 2) {
        tmem_ld32(buf + 32 * (c + 1), rb);
        exp_chunk(ra, c);
        tmem_ld_wait();
        if (c + 2 < 6) tmem_ld32(buf + 32 * (c + 2), ra);
        exp_chunk(rb, c + 1);
        if (c + 2 < 6) tmem_ld_wait();
      }
      s_sum[(t & 3) * 128 + tl] = sum;
      if (t >= 2) mbar_wait(&s_free[grp], (n - 1) & 1);       // never complete p_ready two phases ahead of the epilogue (attention.cuh)
      mbar_arrive(&p_ready[grp]);
      if (p.dbg) w_busy += clock64() - c0;
    }
    if (p.dbg && lane == 0 && quarter == 0) { p.dbg[blockIdx.x * 8 + 1 + 4 * grp] = w_wait; p.dbg[blockIdx.x * 8 + 2 + 4 * grp] = w_busy; }
  } else if (warp < 12) {
    // -------------------------------------------------------------------- epilogue (warps 8..11)
    const int quarter = warp - 8;
    const int tl = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    long long w_wait = 0, w_busy = 0, c0 = 0;
    for (int t = 0; t < T; ++t) {
      const int u = u0 + t, kind = u % 3, pair = u / 3, n = t >> 1, bf = t & 1;
      // the matrix row in this thread's TMEM lane: token `token` of item `item`
      const int item = 2 * pair + (kind == 1 ? 1 : kind == 2 ? (lane >> 4) : 0);
      const int token = kind < 2 ? tl : 128 + quarter * 16 + (lane & 15);
      const int b = item / p.heads, h = item % p.heads;
      if (p.dbg) c0 = clock64();
      mbar_wait(&p_ready[bf], n & 1);
      mbar_wait(&o_full[bf], n & 1);
      if (p.dbg) { w_wait += clock64() - c0; c0 = clock64(); }
      tc_fence_after_sync();
      constexpr int OCH = HD / 16;
      uint32_t o[OCH][16];
#pragma unroll
      for (int qq = 0; qq < OCH; ++qq) tmem_ld16(lane_base + ATT_O_SEP_COL + bf * 64 + 16 * qq, o[qq]);
      tmem_ld_wait();
      const float sum = s_sum[(t & 3) * 128 + tl];
      tc_fence_before_sync();
      mbar_arrive(&s_free[bf]);                             // O and the row sum are in registers: the O columns may be reused
      const float inv = 1.0f / sum;
      uint4* orow = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(b) * ATT_T + token) * p.dim + h * HD);
#pragma unroll
      for (int qq = 0; qq < OCH; ++qq) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(o[qq][8 * v + 0]) * inv, __uint_as_float(o[qq][8 * v + 1]) * inv);
          w.y = pack_bf16(__uint_as_float(o[qq][8 * v + 2]) * inv, __uint_as_float(o[qq][8 * v + 3]) * inv);
          w.z = pack_bf16(__uint_as_float(o[qq][8 * v + 4]) * inv, __uint_as_float(o[qq][8 * v + 5]) * inv);
          w.w = pack_bf16(__uint_as_float(o[qq][8 * v + 6]) * inv, __uint_as_float(o[qq][8 * v + 7]) * inv);
          orow[2 * qq + v] = w;
        }
      }
      if (p.dbg) w_busy += clock64() - c0;
    }
    if (p.dbg && threadIdx.x == 256) { p.dbg[blockIdx.x * 8 + 3] = w_wait; p.dbg[blockIdx.x * 8 + 4] = w_busy; }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  if (p.dbg && threadIdx.x == 0) { p.dbg[blockIdx.x * 8 + 0] = clock64() - t_cta0; p.dbg[blockIdx.x * 8 + 7] = T; }
}

}  // namespace vpb
